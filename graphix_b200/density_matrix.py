"""``B200DensityMatrix`` / ``B200DensityMatrixBackend`` — density-matrix pattern simulation on B200.

Mirrors ``graphix/sim/density_matrix.py:43-496`` (``DensityMatrix``, ``DensityMatrixBackend``): same method
names and semantics.  rho over n qubits is stored *vectorised* as a 2n-qubit register of the gxsv engine
(one row qubit and one column qubit per qubit), so every operation is one of the engine's streaming
kernels instead of the reference's out-of-place ``tensordot``:

* ``U rho U^dagger``          : U on the row qubit, conj(U) on the column qubit — one 2-axis ``k_evolve`` pass
  (diagonal U: two half passes);
* CZ                          : CZ on the two row qubits and on the two column qubits (queued, lazy);
* Kraus channel on k qubits   : ONE 2k-axis pass with the superoperator sum_j |c_j|^2 K_j (x) conj(K_j), where the
  reference runs one full ``evolve`` per Kraus operator (4 for the single-qubit, 16 for the two-qubit
  depolarising channel — the dominant cost of config 5, ``density_matrix.py:416-451``);
* measurement (rank-1 projector |v><v|): contract the row qubit with <v| and the column qubit with |v>, then
  normalise by the trace (``density_matrix.py:367-396`` + ``base_backend.py:450-464``);
* Tr(op_q rho) / Tr(rho)      : a gather over the 2^n diagonal blocks (``k_dm_diag``).
"""
from __future__ import annotations

import ctypes as C
import math
from collections.abc import Iterable, Sequence

import numpy as np

from . import _lib
from .backend import B200StatevectorBackend, NodeIndex
from .mbqc import BasicStates, RandomBranchSelector
from .statevec import B200Statevector, _is_state, _op8


def _ints(xs) -> C.Array:
    return (C.c_int * max(len(xs), 1))(*[int(x) for x in xs])


class B200DensityMatrix:
    """Density matrix in B200 HBM (vectorised).  Constructor follows ``DensityMatrix.__init__``
    (``density_matrix.py:47-123``): ``data`` is a ``State``, an iterable of ``State``, a ``2**n`` vector, a
    ``2**n x 2**n`` matrix, a statevector or a density matrix."""

    def __init__(self, data=BasicStates.PLUS, nqubit: int | None = None, *, max_qubits: int | None = None,
                 device: int | None = None) -> None:
        if nqubit is not None and nqubit < 0:
            raise ValueError("nqubit must be a non-negative integer.")
        self._sv = B200Statevector(nqubit=0, max_qubits=2 * (max_qubits or 0), device=device, fusion=2, strict=True)
        self._sv.set_option(_lib.OPT_DIST, 1)   # no vector norm is imposed: normalisation is by the trace
        self._lib = self._sv._lib
        self._h = self._sv._h
        self.rows: list[int] = []               # engine index of the row qubit of each qubit
        self.cols: list[int] = []
        self._add(data, nqubit)

    # -- construction ------------------------------------------------------------------------------
    def _add(self, data, nqubit: int | None) -> None:
        if isinstance(data, B200DensityMatrix) or (hasattr(data, "rho") and not isinstance(data, np.ndarray)):
            rho = np.asarray(data.rho, dtype=np.complex128)
            self._check_size(rho, nqubit)
            self._add_matrix(rho)
            return
        if isinstance(data, np.ndarray) and data.ndim == 2:
            self._check_matrix(data, nqubit)
            self._add_matrix(np.asarray(data, dtype=np.complex128))
            return
        if not _is_state(data) and isinstance(data, Iterable) and not hasattr(data, "flatten"):
            items = list(data)
            if len(items) != 0 and isinstance(items[0], Iterable) and not _is_state(items[0]):
                rho = np.array([list(r) for r in items], dtype=np.complex128)
                self._check_matrix(rho, nqubit)
                self._add_matrix(rho)
                return
            data = items
        # product of single-qubit states, or a pure-state vector: rho = |psi><psi|
        what, n = B200Statevector._parse_data(data, nqubit)
        if what is None:
            return
        if isinstance(what, np.ndarray):
            self._add_matrix(np.outer(what, what.conj()))
            return
        for s in what:
            v = np.asarray(s.to_statevector_numpy(), dtype=np.complex128)
            self._add_pure_qubit(complex(v[0]), complex(v[1]))

    @staticmethod
    def _check_size(rho: np.ndarray, nqubit: int | None) -> None:
        if nqubit is not None and rho.shape != (2 ** nqubit, 2 ** nqubit):
            raise ValueError(
                f"Inconsistent parameters between nqubit = {nqubit} and the shape of the provided density matrix = {rho.shape}."
            )

    def _check_matrix(self, rho: np.ndarray, nqubit: int | None) -> None:
        d = rho.shape[0]
        if rho.ndim != 2 or rho.shape[0] != rho.shape[1] or d & (d - 1):
            raise ValueError("Cannot interpret the provided density matrix as a qubit operator.")
        self._check_size(rho, nqubit)
        if not np.isclose(np.trace(rho), 1.0):
            raise ValueError("Density matrix must have unit trace.")
        if not np.allclose(rho, rho.conj().T) or np.min(np.linalg.eigvalsh((rho + rho.conj().T) / 2)) < -1e-9:
            raise ValueError("Density matrix must be positive semi-definite.")

    def _check(self, rc: int) -> None:
        if rc != _lib.GXSV_OK:
            _lib.raise_for(self._lib, self._h, rc)

    def _add_pure_qubit(self, s0: complex, s1: complex) -> None:
        base = self._sv.nqubit
        for a, b in ((s0, s1), (s0.conjugate(), s1.conjugate())):
            self._check(self._lib.gxsv_add_qubit(self._h, (C.c_double * 4)(a.real, a.imag, b.real, b.imag)))
        self.rows.append(base)
        self.cols.append(base + 1)

    def _add_matrix(self, rho: np.ndarray) -> None:
        k = rho.shape[0].bit_length() - 1
        if k == 0:
            return
        base = self._sv.nqubit
        buf = np.ascontiguousarray(rho, dtype=np.complex128).reshape(-1)   # index = row * 2^k + col
        self._check(self._lib.gxsv_tensor(self._h, 2 * k, buf.ctypes.data_as(C.c_void_p)))
        self.rows.extend(range(base, base + k))
        self.cols.extend(range(base + k, base + 2 * k))

    # -- reference surface -------------------------------------------------------------------------------
    @property
    def nqubit(self) -> int:
        return len(self.rows)

    def dims(self) -> tuple[int, int]:
        return (1 << self.nqubit, 1 << self.nqubit)

    def _canonical(self) -> None:
        """Relabel the engine so that its qubit order is [row qubits..., column qubits...] (free)."""
        want = self.rows + self.cols
        if want != list(range(len(want))):
            self._check(self._lib.gxsv_permute(self._h, _ints(want), len(want)))
            n = self.nqubit
            self.rows = list(range(n))
            self.cols = list(range(n, 2 * n))

    @property
    def rho(self) -> np.ndarray:
        n = self.nqubit
        self._canonical()
        return self._sv.flatten().reshape(1 << n, 1 << n)

    def flatten(self) -> np.ndarray:
        """density_matrix.py:412-414."""
        return self.rho.reshape(-1)

    def add_nodes(self, nqubit: int, data) -> None:
        """density_matrix.py:166-193."""
        self._add(data, nqubit)

    def tensor(self, other) -> None:
        """density_matrix.py:296-310."""
        self._add(other, None)

    def entangle(self, qubits: tuple[int, int]) -> None:
        """CZ rho CZ — density_matrix.py:344-352."""
        a, b = self._q(qubits[0]), self._q(qubits[1])
        self._check(self._lib.gxsv_entangle(self._h, self.rows[a], self.rows[b]))
        self._check(self._lib.gxsv_entangle(self._h, self.cols[a], self.cols[b]))

    def _q(self, q: int) -> int:
        if not 0 <= q < self.nqubit:
            raise IndexError(f"Qubit index {q} out of range [0, {self.nqubit - 1}]")
        return int(q)

    def evolve_single(self, op, qubit: int) -> None:
        """op rho op^dagger on one qubit — density_matrix.py:195-213."""
        q = self._q(qubit)
        m = np.asarray(op, dtype=np.complex128)
        if m.shape != (2, 2):
            raise ValueError("op must be 2*2 matrix.")
        if m[0, 1] == 0 and m[1, 0] == 0:
            self._check(self._lib.gxsv_evolve_single(self._h, self.rows[q], _op8(m)))
            self._check(self._lib.gxsv_evolve_single(self._h, self.cols[q], _op8(m.conj())))
            return
        self._super(np.kron(m, m.conj()), [q])

    def evolve(self, op, qubits: Sequence[int]) -> None:
        """density_matrix.py:215-261."""
        m = np.asarray(op, dtype=np.complex128)
        k = len(qubits)
        if m.shape != (1 << k, 1 << k):
            raise ValueError("The dimension of the operator doesn't match the number of targets.")
        if len(set(qubits)) != k:
            raise ValueError("A repeated target qubit index is not possible.")
        self._super(np.kron(m, m.conj()), [self._q(q) for q in qubits])

    def _super(self, sop: np.ndarray, qubits: Sequence[int]) -> None:
        """One pass with a superoperator whose index is (row bits of `qubits`..., column bits of `qubits`...)."""
        k = len(qubits)
        if 2 * k > 4:
            raise NotImplementedError("operators on more than two qubits are not supported by the density-matrix path")
        targets = [self.rows[q] for q in qubits] + [self.cols[q] for q in qubits]
        buf = np.ascontiguousarray(sop, dtype=np.complex128)
        self._check(self._lib.gxsv_evolve(self._h, 2 * k, _ints(targets), buf.ctypes.data_as(C.c_void_p)))

    def cnot(self, edge: tuple[int, int]) -> None:
        self.evolve(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), edge)

    def swap(self, qubits: tuple[int, int]) -> None:
        """density_matrix.py:325-331 (a relabel here)."""
        a, b = self._q(qubits[0]), self._q(qubits[1])
        self.rows[a], self.rows[b] = self.rows[b], self.rows[a]
        self.cols[a], self.cols[b] = self.cols[b], self.cols[a]

    def permute(self, permutation: Sequence[int]) -> None:
        """density_matrix.py:333-342."""
        n = self.nqubit
        if len(permutation) != n:
            raise ValueError(f"Permutation has length {len(permutation)}, but {n} qubits expected.")
        if set(permutation) != set(range(n)):
            raise ValueError(f"{permutation} is not a permutation.")
        self.rows = [self.rows[p] for p in permutation]
        self.cols = [self.cols[p] for p in permutation]

    def _traces(self, target: int, op) -> tuple[complex, complex]:
        out = (C.c_double * 4)()
        m8 = _op8(op) if op is not None else (C.c_double * 8)()
        self._check(self._lib.gxsv_dm_trace_expect(self._h, self.nqubit, _ints(self.rows), _ints(self.cols), target, m8, out))
        return complex(out[0], out[1]), complex(out[2], out[3])

    def expectation_single(self, op, qubit: int) -> complex:
        """Tr(op_q rho) / Tr(rho) — density_matrix.py:263-290."""
        if not 0 <= qubit < self.nqubit:
            raise ValueError(f"Wrong target qubit {qubit}. Must between 0 and {self.nqubit - 1}.")
        m = np.asarray(op, dtype=np.complex128)
        if m.shape != (2, 2):
            raise ValueError("op must be 2x2 matrix.")
        num, tr = self._traces(int(qubit), m)
        return num / tr

    def trace(self) -> complex:
        """Tr(rho) (1 after every projection); host read-out, small registers only."""
        return complex(np.trace(self.rho))

    def project_qubit(self, op, qubit: int) -> None:
        """op rho op^dagger, trace the qubit out, normalise — base_backend.py:450-464 with
        density_matrix.py:367-396, for the rank-1 projectors that ``measure`` builds."""
        q = self._q(qubit)
        m = np.asarray(op, dtype=np.complex128)
        if abs(m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) > 1e-13 * float(np.sum(np.abs(m) ** 2)):
            raise NotImplementedError("the density-matrix path projects rank-1 operators only")
        r, c = self.rows[q], self.cols[q]
        # <v| on the row qubit (rows of P = |v><v| are proportional to <v|), |v> on the column qubit (rows of conj(P))
        self._check(self._lib.gxsv_project_qubit(self._h, r, _op8(m), 0.0, None))
        c -= 1 if c > r else 0
        self._check(self._lib.gxsv_project_qubit(self._h, c, _op8(m.conj()), 0.0, None))
        removed = sorted((self.rows[q], self.cols[q]))
        del self.rows[q], self.cols[q]

        def shift(i: int) -> int:
            return i - (i > removed[0]) - (i > removed[1])

        self.rows = [shift(i) for i in self.rows]
        self.cols = [shift(i) for i in self.cols]
        self.normalize()

    def remove_qubit(self, qubit: int) -> None:
        """Trace a qubit out and normalise — density_matrix.py:367-371 (``ptrace`` + ``normalize``).

        Done with the kernels already there: one 2-axis pass moves rho_00 + rho_11 of the qubit into its |0><0| block
        (superoperator with S[00,00] = S[00,11] = 1), then the |0><0| projection removes the qubit."""
        q = self._q(qubit)
        sop = np.zeros((4, 4), dtype=np.complex128)
        sop[0, 0] = sop[0, 3] = 1.0
        self._super(sop, [q])
        self.project_qubit(np.array([[1, 0], [0, 0]], dtype=np.complex128), q)

    def ptrace(self, qargs) -> None:
        """density_matrix.py:373-396; the result is also normalised (the engine keeps rho at unit trace)."""
        qs = [qargs] if isinstance(qargs, int) else sorted(qargs, reverse=True)
        for q in qs:
            self.remove_qubit(q)

    def normalize(self) -> None:
        """rho /= Tr(rho) — density_matrix.py:354-365."""
        if self.nqubit == 0:
            tr = complex(self._sv.flatten()[0])
            if tr != 0:
                self._check(self._lib.gxsv_scale_host(self._h, (1 / tr).real, (1 / tr).imag))
            return
        _, tr = self._traces(-1, None)       # stored units
        t = tr.real
        if t == 0.0 or math.isnan(t):
            raise RuntimeError("Attempted to normalize a density matrix with zero trace.")
        # engine: true = g * h * stored; choose g = 1/|Tr(stored)| and put the remaining sign/phase into h
        self._check(self._lib.gxsv_set_norm(self._h, abs(tr) ** 2))
        ph = abs(tr) / tr
        # the host scalar may still carry factors from earlier gates (e.g. m00 of diagonal operators): reset it
        self._reset_hscale(ph)

    def _reset_hscale(self, value: complex) -> None:
        h = C.c_double * 2
        cur = h()
        self._check(self._lib.gxsv_get_hscale(self._h, cur))
        c = complex(cur[0], cur[1])
        f = value / c if c != 0 else value
        self._check(self._lib.gxsv_scale_host(self._h, f.real, f.imag))

    def apply_channel(self, channel, qargs: Sequence[int]) -> None:
        """rho <- sum_j |c_j|^2 K_j rho K_j^dagger — density_matrix.py:416-451, as ONE superoperator pass."""
        ops = list(channel)
        if not ops:
            raise TypeError("Can't apply a channel that is not a Channel object.")
        k = len(qargs)
        dim = 1 << k
        sop = np.zeros((dim * dim, dim * dim), dtype=np.complex128)
        check = np.zeros((dim, dim), dtype=np.complex128)
        for kd in ops:
            op = np.asarray(kd.operator, dtype=np.complex128)
            if op.shape != (dim, dim):
                raise ValueError("The dimension of the operator doesn't match the number of targets.")
            w = abs(kd.coef) ** 2
            sop += w * np.kron(op, op.conj())
            check += w * (op.conj().T @ op)
        if not np.allclose(check, np.eye(dim)):
            raise ValueError("The output density matrix is not normalized, check the channel definition.")
        self._super(sop, [self._q(q) for q in qargs])

    def apply_noise(self, qubits: Sequence[int], noise) -> None:
        """density_matrix.py:453-465."""
        self.apply_channel(noise.to_krauschannel(), qubits)

    def fidelity(self, statevec) -> float:
        """<psi| rho |psi> — density_matrix.py:398-410."""
        psi = np.asarray(statevec.flatten(), dtype=np.complex128)
        return float(np.vdot(psi, self.rho @ psi).real)

    def __str__(self) -> str:
        return f"B200DensityMatrix(nqubit={self.nqubit})"

    __repr__ = __str__

    def stats(self):
        return self._sv.stats()

    def reset_stats(self) -> None:
        self._sv.reset_stats()

    def set_option(self, key: int, value: int) -> None:
        self._sv.set_option(key, value)

    def sync(self) -> None:
        self._sv.sync()

    def reset(self) -> None:
        self._sv.reset()
        self.rows, self.cols = [], []


class B200DensityMatrixBackend(B200StatevectorBackend):
    """``DensityMatrixBackend`` (density_matrix.py:492-496) on B200: same ``Backend`` contract, supports noise."""

    def __init__(self, state: B200DensityMatrix | None = None, *, node_index=None, branch_selector=None,
                 symbolic: bool = False, device: int | None = None, max_qubits: int | None = None) -> None:
        if symbolic:
            raise ValueError("the B200 density-matrix backend does not support `symbolic` simulation.")
        if state is None:
            state = B200DensityMatrix(nqubit=0, max_qubits=max_qubits, device=device)
        set_ = object.__setattr__
        set_(self, "state", state)
        set_(self, "node_index", NodeIndex() if node_index is None else node_index)
        set_(self, "branch_selector", RandomBranchSelector() if branch_selector is None else branch_selector)
        set_(self, "symbolic", False)

    @classmethod
    def with_capacity(cls, max_qubits: int, state=None, **kwargs) -> "B200DensityMatrixBackend":
        dev = kwargs.pop("device", None)
        if state is None:
            state = B200DensityMatrix(nqubit=0, max_qubits=max_qubits, device=dev)
        return cls(state, **kwargs)

    def measure(self, node: int, measurement, rng=None, *, stacklevel: int = 1):
        """base_backend.py:812-856."""
        loc = self.node_index.index(node)
        bloch = measurement.to_bloch()
        x, y, z = (float(v) for v in bloch.plane.polar(bloch.angle))

        def proj(sign: float) -> np.ndarray:
            return np.array([[0.5 + 0.5 * sign * z, 0.5 * sign * (x - 1j * y)],
                             [0.5 * sign * (x + 1j * y), 0.5 - 0.5 * sign * z]], dtype=np.complex128)

        def f_expectation0() -> float:
            v = self.state.expectation_single(proj(1.0), loc)
            assert abs(v.imag) <= 1e-10
            return v.real

        outcome = self.branch_selector.measure(node, f_expectation0, rng, stacklevel=stacklevel + 1)
        self.state.project_qubit(proj(-1.0 if outcome else 1.0), loc)
        self.node_index.remove(node)
        return outcome

    def apply_noise(self, cmd) -> None:
        """base_backend.py:864-874."""
        indices = [self.node_index.index(i) for i in cmd.nodes]
        self.state.apply_noise(indices, cmd.noise)


def _register_with_graphix() -> None:
    """When graphix is importable, make ``isinstance(state, graphix.sim.density_matrix.DensityMatrix)`` (and ``DenseState``)
    hold for ``B200DensityMatrix`` — the reference's helpers dispatch on that class (tests/test_pattern.py:39-44).  A
    virtual-subclass registration: no graphix code is inherited."""
    try:
        from graphix.sim.base_backend import DenseState  # noqa: PLC0415
        from graphix.sim.density_matrix import DensityMatrix  # noqa: PLC0415
    except Exception:  # noqa: BLE001 - graphix absent
        return
    DenseState.register(B200DensityMatrix)
    DensityMatrix.register(B200DensityMatrix)


_register_with_graphix()
