"""Pattern driver mirroring ``graphix.simulator.PatternSimulator`` (simulator.py:306-479).

With graphix installed, use graphix's own ``PatternSimulator`` /
``pattern.simulate(backend=B200StatevectorBackend(...))``; this module is the
same loop for environments without graphix (it walks ``graphix_b200.mbqc``
commands, or any objects with the same ``kind.name`` / field names).
"""
from __future__ import annotations

from collections.abc import Mapping

from .backend import B200StatevectorBackend
from .mbqc import BasicStates, BranchSelector, RandomBranchSelector


class DefaultMeasureMethod:
    """Signal-adaptive measurement — simulator.py:180-257."""

    def __init__(self, results: Mapping[int, int] | None = None) -> None:
        self.results: dict[int, int] = {} if results is None else dict(results)

    def describe_measurement(self, cmd):
        """simulator.py:205-228: s-signal applies Clifford X, t-signal Clifford Z."""
        res = self.results
        s_signal = sum(res[j] for j in cmd.s_domain) & 1
        t_signal = sum(res[j] for j in cmd.t_domain) & 1
        m = cmd.measurement
        if s_signal:
            m = m.clifford_x() if hasattr(m, "clifford_x") else m.clifford(_ref_clifford("X"))
        if t_signal:
            m = m.clifford_z() if hasattr(m, "clifford_z") else m.clifford(_ref_clifford("Z"))
        return m

    def measure(self, backend, cmd, noise_model=None, rng=None, *, stacklevel: int = 1) -> None:
        """simulator.py:94-124."""
        description = self.describe_measurement(cmd)
        result = backend.measure(cmd.node, description, rng=rng, stacklevel=stacklevel + 1)
        if noise_model is not None:
            result = noise_model.confuse_result(cmd, result, rng=rng, stacklevel=stacklevel + 1)
        self.store_measurement_outcome(cmd.node, result)

    def measurement_outcome(self, node: int) -> int:
        return self.results[node]

    def store_measurement_outcome(self, node: int, result: int) -> None:
        self.results[node] = result

    def check_domain(self, domain) -> bool:
        """simulator.py:171-179."""
        res = self.results
        return sum(res[j] for j in domain) % 2 == 1


def _ref_clifford(name: str):
    from graphix.clifford import Clifford  # noqa: PLC0415 - only reached with graphix command objects

    return getattr(Clifford, name)


class PatternSimulator:
    """simulator.py:306-479.

    ``backend`` is a backend instance, or ``"b200"`` / ``"statevector"`` to build a
    ``B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=...)``
    the way ``_initialize_backend`` does (simulator.py:568-576).
    """

    def __init__(self, pattern, backend="b200", *, branch_selector: BranchSelector | None = None,
                 measure_method=None, noise_model=None, symbolic: bool = False, stacklevel: int = 1,
                 **backend_kwargs) -> None:
        if isinstance(backend, str):
            if backend in ("densitymatrix", "b200-densitymatrix"):
                from .density_matrix import B200DensityMatrixBackend  # noqa: PLC0415

                selector = RandomBranchSelector() if branch_selector is None else branch_selector
                backend = B200DensityMatrixBackend.with_capacity(pattern.max_space(), branch_selector=selector,
                                                                 **backend_kwargs)
                self.backend = backend
                self.noise_model = noise_model
                self._pattern = pattern
                self._measure_method = measure_method or DefaultMeasureMethod()
                return
            if backend not in ("b200", "statevector"):
                raise ValueError(f"Unknown backend {backend}.")
            if noise_model is not None:
                raise ValueError("`noise_model` cannot be specified for state vector backend.")
            if symbolic:
                raise ValueError(
                    "Statevector backend does not support `symbolic` simulation. Consider using backend in `graphix-symbolic` plugin."
                )
            selector = RandomBranchSelector() if branch_selector is None else branch_selector
            backend = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=selector,
                                                           **backend_kwargs)
        else:
            if branch_selector is not None:
                raise ValueError("`branch_selector` cannot be specified if `backend` is already instantiated.")
            if symbolic:
                raise ValueError("`symbolic` cannot be specified if `backend` is already instantiated.")
        self.backend = backend
        self.noise_model = noise_model
        self._pattern = pattern
        self._measure_method = measure_method or DefaultMeasureMethod()

    @property
    def pattern(self):
        return self._pattern

    @property
    def measure_method(self):
        return self._measure_method

    def run(self, input_state=BasicStates.PLUS, rng=None, *, stacklevel: int = 1) -> None:
        """simulator.py:396-479."""
        backend = self.backend
        pattern = self._pattern
        initial_nqubit = backend.nqubit
        if input_state is None:
            n_in = len(pattern.input_nodes)
            if initial_nqubit != n_in:
                raise ValueError(
                    f"`input_state` is `None`: the backend is expected to have {n_in} input nodes already prepared, but {initial_nqubit} were found."
                )
        elif initial_nqubit != 0:
            raise ValueError(
                f"`input_state` is not `None`: the backend is expected to have no pre-allocated qubits, but has {initial_nqubit} qubits."
            )
        if input_state is not None:
            backend.add_nodes(pattern.input_nodes, input_state)
        if self.noise_model is None:
            sequence = pattern
        else:
            # simulator.py:437-441: noise on the input nodes, then the noise-transpiled command list
            sequence = self.noise_model.input_nodes(pattern.input_nodes, rng=rng) if input_state is not None else []
            sequence.extend(self.noise_model.transpile(pattern, rng=rng))
        pattern.check_runnability()
        mm = self._measure_method
        noise_model = self.noise_model
        for cmd in sequence:
            kind = cmd.kind.name
            if kind == "N":
                backend.add_nodes(nodes=[cmd.node], data=cmd.state)
            elif kind == "E":
                backend.entangle_nodes(edge=cmd.nodes)
            elif kind == "M":
                mm.measure(backend, cmd, noise_model=noise_model, rng=rng, stacklevel=stacklevel + 1)
            elif kind in ("X", "Z"):
                if mm.check_domain(cmd.domain):
                    backend.correct_byproduct(cmd)
            elif kind == "C":
                backend.apply_clifford(cmd.node, cmd.clifford)
            elif kind == "T":
                pass
            elif kind == "ApplyNoise":
                if cmd.domain is None or mm.check_domain(cmd.domain):
                    backend.apply_noise(cmd)
            elif kind == "S":
                raise ValueError("S commands unexpected in simulated patterns.")
            else:
                raise ValueError(f"unsupported command kind {kind}")
        backend.finalize(output_nodes=pattern.output_nodes)
