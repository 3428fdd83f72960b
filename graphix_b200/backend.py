"""``B200StatevectorBackend`` — the drop-in ``Backend`` for graphix pattern simulation.

Boundary reproduced: ``graphix.sim.base_backend.Backend``
(base_backend.py:545-725) as implemented by ``DenseStateBackend``
(base_backend.py:739-898) and ``StatevectorBackend`` (statevec.py:765-832).
``PatternSimulator.run`` (simulator.py:396-479) calls exactly ``nqubit``,
``add_nodes``, ``entangle_nodes``, ``measure``, ``correct_byproduct``,
``apply_clifford``, ``apply_noise``, ``finalize`` and reads ``state``.

When graphix is importable the class also inherits from its
``DenseStateBackend`` so that ``isinstance(backend, Backend)``
(simulator.py:549) and ``isinstance(backend, DenseStateBackend)``
(transpiler.py:1165) hold; the host logic below is the same either way and only
duck-types the graphix objects it is handed (``Measurement.to_bloch()``,
``plane.polar``, ``cmd.kind.name``, ``clifford.matrix``, ``State.to_statevector_numpy``).
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Iterable, Iterator, Sequence

from . import _lib
from .mbqc import BasicStates, BranchSelector, Ops, RandomBranchSelector
from .statevec import B200Statevector, NoiseNotSupportedError


class NodeIndex:
    """Node label <-> qubit index — graphix/sim/base_backend.py:216-321."""

    def __init__(self) -> None:
        self._list: list[int] = []
        self._dict: dict[int, int] = {}

    def __getitem__(self, index: int) -> int:
        return self._list[index]

    def index(self, node: int) -> int:
        return self._dict[node]

    def __iter__(self) -> Iterator[int]:
        return iter(self._list)

    def __len__(self) -> int:
        return len(self._list)

    def extend(self, nodes: Iterable[int]) -> None:
        base = len(self._list)
        self._list.extend(nodes)
        for i in range(base, len(self._list)):
            self._dict[self._list[i]] = i

    def remove(self, node: int) -> None:
        index = self._dict.pop(node)
        del self._list[index]
        for i in range(index, len(self._list)):
            self._dict[self._list[i]] = i

    def swap(self, i: int, j: int) -> None:
        a, b = self._list[i], self._list[j]
        self._list[i], self._list[j] = b, a
        self._dict[a], self._dict[b] = j, i

    def clear(self) -> None:
        self._list = []
        self._dict = {}


def _graphix_base():
    try:
        from graphix.sim.base_backend import DenseStateBackend  # noqa: PLC0415

        return DenseStateBackend
    except Exception:  # noqa: BLE001 - graphix absent (GPU box) or not importable
        return object


_Base = _graphix_base()


class B200StatevectorBackend(_Base):  # type: ignore[misc, valid-type]
    """MBQC statevector backend whose state lives in B200 HBM.

    Parameters (keyword-only like ``DenseStateBackendKwargs``, base_backend.py:731-736)
    ----------
    state : B200Statevector, optional
        Initial state; default is the 0-qubit state.
    node_index : NodeIndex, optional
    branch_selector : BranchSelector, optional
        Default ``RandomBranchSelector()``.
    symbolic : bool
        Must be False (statevec.py:802-805).
    """

    def __init__(self, state: B200Statevector | None = None, *, node_index=None, branch_selector=None,
                 symbolic: bool = False, device: int | None = None, fusion: int = 2, strict: bool = True,
                 batch: int | None = None) -> None:
        if symbolic:
            raise ValueError(
                "Statevector backend does not support `symbolic` simulation. Consider using backend in `graphix-symbolic` plugin."
            )
        if state is None:
            state = B200Statevector(nqubit=0, device=device, fusion=fusion, strict=strict, batch=batch)
        elif not isinstance(state, B200Statevector) and not hasattr(state, "project_qubit"):
            state = B200Statevector(state, device=device, fusion=fusion, strict=strict, batch=batch)
        # (any other object with the DenseState surface — project_qubit, expectation_single, ... — is used as is
        #  through the generic calls below; this is how the host logic is exercised without a GPU in the tests)
        set_ = object.__setattr__  # the graphix base is a frozen dataclass
        set_(self, "state", state)
        set_(self, "node_index", NodeIndex() if node_index is None else node_index)
        set_(self, "branch_selector", RandomBranchSelector() if branch_selector is None else branch_selector)
        set_(self, "symbolic", False)
        set_(self, "_lib", getattr(state, "_lib", None))
        set_(self, "_op", (C.c_double * 8)())

    @classmethod
    def with_capacity(cls, max_qubits: int, state=None, **kwargs) -> "B200StatevectorBackend":
        """Preallocate ``2**max_qubits`` amplitudes — statevec.py:807-832."""
        dev = kwargs.pop("device", None)
        fusion = kwargs.pop("fusion", 2)
        strict = kwargs.pop("strict", True)
        batch = kwargs.pop("batch", None)
        if state is None:
            st = B200Statevector(nqubit=0, max_qubits=max_qubits, device=dev, fusion=fusion, strict=strict, batch=batch)
        else:
            st = B200Statevector(state, max_qubits=max_qubits, device=dev, fusion=fusion, strict=strict, batch=batch)
        return cls(st, **kwargs)

    # -- Backend contract -------------------------------------------------------
    @property
    def nqubit(self) -> int:
        """base_backend.py:894-898."""
        return self.state.nqubit

    def add_nodes(self, nodes: Sequence[int], data=BasicStates.PLUS) -> None:
        """base_backend.py:778-797."""
        self.state.add_nodes(nqubit=len(nodes), data=data)
        self.node_index.extend(nodes)

    def entangle_nodes(self, edge: tuple[int, int]) -> None:
        """base_backend.py:799-810."""
        target = self.node_index.index(edge[0])
        control = self.node_index.index(edge[1])
        self.state.entangle((target, control))

    def measure(self, node: int, measurement, rng=None, *, stacklevel: int = 1):
        """Measure ``node`` and remove its qubit — base_backend.py:812-856.

        The projector is 1/2 (I + (-1)^r v.sigma) for the Bloch vector v of the
        (already signal-adapted) measurement — base_backend.py:506-542.
        """
        loc = self.node_index.index(node)
        bloch = measurement.to_bloch()
        x, y, z = bloch.plane.polar(bloch.angle)
        x, y, z = float(x), float(y), float(z)
        state = self.state
        op = self._op

        def fill(sign: float) -> None:
            op[0] = 0.5 + 0.5 * sign * z
            op[1] = 0.0
            op[2] = 0.5 * sign * x
            op[3] = -0.5 * sign * y
            op[4] = 0.5 * sign * x
            op[5] = 0.5 * sign * y
            op[6] = 0.5 - 0.5 * sign * z
            op[7] = 0.0

        def as_matrix():
            import numpy as np  # noqa: PLC0415

            return np.array([[complex(op[0], op[1]), complex(op[2], op[3])],
                             [complex(op[4], op[5]), complex(op[6], op[7])]])

        def f_expectation0() -> float:
            fill(1.0)
            if self._lib is None:
                v = state.expectation_single(as_matrix(), loc)
                assert abs(v.imag) <= 1e-10
                return v.real
            out = (C.c_double * 2)()
            state._check(self._lib.gxsv_expectation_single(state._h, loc, op, out))
            assert abs(out[1]) <= 1e-10
            return out[0]

        outcome = self.branch_selector.measure(node, f_expectation0, rng, stacklevel=stacklevel + 1)
        fill(-1.0 if outcome else 1.0)
        if self._lib is None:
            state.project_qubit(as_matrix(), loc)
        else:
            state._check(self._lib.gxsv_project_qubit(state._h, loc, op, 1e-10, None))
        self.node_index.remove(node)
        return outcome

    def correct_byproduct(self, cmd) -> None:
        """base_backend.py:858-862."""
        op = Ops.X if cmd.kind.name == "X" else Ops.Z
        self.apply_single(node=cmd.node, op=op)

    def apply_noise(self, cmd) -> None:  # noqa: ARG002
        """base_backend.py:864-874 -> DenseState.apply_noise default (:487-503)."""
        raise NoiseNotSupportedError

    def apply_single(self, node: int, op) -> None:
        """base_backend.py:876-879."""
        self.state.evolve_single(op=op, qubit=self.node_index.index(node))

    def apply_clifford(self, node: int, clifford) -> None:
        """base_backend.py:881-885."""
        self.state.evolve_single(clifford.matrix, self.node_index.index(node))

    def finalize(self, output_nodes: Sequence[int]) -> None:
        """base_backend.py:887-892."""
        self.state.permute([self.node_index.index(node) for node in output_nodes])
        self.node_index.clear()
        self.node_index.extend(output_nodes)
        sync = getattr(self.state, "sync", None)
        if sync is not None:
            sync()    # non-strict mode: deferred zero-norm errors of the run surface here at the latest

    def __repr__(self) -> str:
        return f"B200StatevectorBackend(state={self.state!r}, branch_selector={self.branch_selector!r})"


__all__ = ["B200StatevectorBackend", "NodeIndex", "BranchSelector", "_lib"]
