"""CPU oracle: restatement of the reference statevector backend.

TEST INFRASTRUCTURE ONLY — importable from ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never from
``graphix_b200``.  Parity is PINNED: ``tests/test_oracle.py`` checks it against
golden outputs produced by importing the reference itself
(``tests/golden/make_golden.py``).

``OracleStatevector`` restates ``graphix/sim/statevec.py:52-750`` on top of the C
kernels of ``statevec_oracle.c`` (one per numba kernel, ``statevec.py:841-1198``);
``OracleBackend`` restates ``DenseStateBackend`` (``base_backend.py:739-898``)
including ``_outcome_to_operator_matrix`` (``base_backend.py:506-542``).
It speaks the vocabulary of ``graphix_b200.mbqc`` (or graphix's own objects).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from collections.abc import Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

_lib = None
_c128p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "statevec_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_tensor.argtypes = [_c128p, _c128p, C.c_int, C.c_int]
        L.orc_add_default_node.argtypes = [_c128p, C.c_int]
        L.orc_swap.argtypes = [_c128p, C.c_int, C.c_int, C.c_int]
        L.orc_entangle.argtypes = [_c128p, C.c_int, C.c_int, C.c_int]
        L.orc_evolve_single.argtypes = [_c128p, _c128p, C.c_int, C.c_int]
        L.orc_expectation_single.argtypes = [_c128p, _c128p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_project_qubit.argtypes = [_c128p, _c128p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.orc_project_qubit.restype = C.c_int
        for f in (L.orc_tensor, L.orc_add_default_node, L.orc_swap, L.orc_entangle, L.orc_evolve_single,
                  L.orc_expectation_single):
            f.restype = None
        _lib = L
    return _lib


def _op(op) -> np.ndarray:
    a = np.asarray(op)
    if a.dtype == np.object_:
        raise TypeError("Statevector backend does not support symbolic operators.")
    return np.ascontiguousarray(a, dtype=np.complex128).reshape(4)


def _is_state(x) -> bool:
    return hasattr(x, "to_statevector_numpy")


class OracleStatevector:
    """statevec.py:52-750 (flat c128 array with capacity; qubit 0 = MSB)."""

    def __init__(self, data=None, nqubit: int | None = None, max_qubits: int | None = None) -> None:
        # statevec.py:93-216
        if data is None:
            data = [] if nqubit in (None, 0) else _plus()
        if isinstance(data, OracleStatevector) or (hasattr(data, "flatten") and hasattr(data, "nqubit")
                                                   and not isinstance(data, np.ndarray)):
            psi = np.array(data.flatten(), dtype=np.complex128)
            n = int(data.nqubit)
            if nqubit is not None and len(psi) != 1 << nqubit:
                raise ValueError("Inconsistent parameters between nqubit and the input state.")
        elif _is_state(data):
            n = 1 if nqubit is None else nqubit
            psi = _kron([data] * n)
        else:
            items = list(data)
            if len(items) == 0:
                if nqubit not in (None, 0):
                    raise ValueError("`nqubit` is not null but input state is empty.")
                n, psi = 0, np.array([1], dtype=np.complex128)
            elif _is_state(items[0]):
                if nqubit is not None and nqubit != len(items):
                    raise ValueError("Mismatch between nqubit and length of input state.")
                n, psi = len(items), _kron(items)
            else:
                length = len(items)
                n = length.bit_length() - 1
                if nqubit is None:
                    if length & (length - 1):
                        raise ValueError(f"Length of input data is not a power of two: {length}.")
                elif nqubit != n:
                    raise ValueError("Mismatch between nqubit and inferred nqubit.")
                psi = np.array(items, dtype=np.complex128)
                if not np.isclose(np.linalg.norm(psi), 1.0):
                    raise ValueError("Input state is not normalized.")
        if max_qubits is not None and max_qubits < n:
            raise ValueError("`max_qubits` can't be smaller than the length of input state.")
        self._psi = np.ascontiguousarray(psi)
        self._nqubit = n
        self._max_qubits = n
        self.ensure_capacity(n if max_qubits is None else max_qubits)

    @property
    def nqubit(self) -> int:
        return self._nqubit

    @property
    def max_qubits(self) -> int:
        return self._max_qubits

    @property
    def psi(self) -> np.ndarray:
        return self._psi[: 1 << self._nqubit]

    def flatten(self) -> np.ndarray:
        return self.psi

    def ensure_capacity(self, required_qubits: int) -> None:  # statevec.py:243-259
        if required_qubits > self._max_qubits or len(self._psi) < (1 << required_qubits):
            new = np.empty(1 << required_qubits, dtype=np.complex128)
            new[: len(self._psi)] = self._psi
            self._psi = new
            self._max_qubits = required_qubits

    def _check_bounds(self, q: int) -> None:  # statevec.py:526-542
        if not 0 <= q < self._nqubit:
            raise IndexError(f"Qubit index {q} out of range [0, {self._nqubit - 1}]")

    def add_nodes(self, nqubit: int, data) -> None:  # statevec.py:269-300
        if nqubit == 1 and _is_plus(data):
            self.ensure_capacity(self._nqubit + 1)
            lib().orc_add_default_node(self._psi, self._nqubit)
            self._nqubit += 1
        else:
            self.tensor(OracleStatevector(data, nqubit=nqubit))

    def tensor(self, other) -> None:  # statevec.py:508-524
        o = np.ascontiguousarray(other.flatten(), dtype=np.complex128)
        self.ensure_capacity(self._nqubit + other.nqubit)
        lib().orc_tensor(self._psi, o, self._nqubit, other.nqubit)
        self._nqubit += other.nqubit

    def entangle(self, qubits) -> None:  # statevec.py:302-315
        for q in qubits:
            self._check_bounds(q)
        lib().orc_entangle(self._psi, self._nqubit, int(qubits[0]), int(qubits[1]))

    def evolve_single(self, op, qubit: int) -> None:  # statevec.py:317-332
        self._check_bounds(qubit)
        lib().orc_evolve_single(self._psi, _op(op), self._nqubit, int(qubit))

    def evolve(self, op, qubits: Sequence[int]) -> None:  # statevec.py:361-395 (numpy einsum fallback there too)
        nq = len(qubits)
        n = self._nqubit
        op_t = np.asarray(op, dtype=np.complex128).reshape((2,) * (2 * nq))
        psi_t = self.psi.reshape((2,) * n)
        out_idx = list(range(n, n + nq))
        res_idx = list(range(n))
        for i, q in enumerate(qubits):
            res_idx[q] = out_idx[i]
        self._psi[: 1 << n] = np.einsum(op_t, out_idx + list(qubits), psi_t, list(range(n)), res_idx).reshape(1 << n)

    def expectation_value(self, op, qubits: Sequence[int]) -> complex:  # statevec.py:397-415
        sv = OracleStatevector(self.flatten().copy())
        sv.evolve(op, qubits)
        return complex(np.dot(self.flatten().conjugate(), sv.flatten()))

    def expectation_single(self, op, qubit: int) -> complex:  # statevec.py:334-359
        self._check_bounds(qubit)
        out = (C.c_double * 2)()
        lib().orc_expectation_single(self._psi, _op(op), self._nqubit, int(qubit), out)
        return complex(out[0], out[1])

    def project_qubit(self, op, qubit: int) -> None:  # statevec.py:430-495
        self._check_bounds(qubit)
        n = lib().orc_project_qubit(self._psi, _op(op), self._nqubit, int(qubit), 1e-10, None)
        if n < 0:
            raise RuntimeError(f"Attempted to remove qubit {qubit} from 0-norm statevector.")
        self._nqubit = n

    def swap(self, qubits) -> None:  # statevec.py:497-506
        lib().orc_swap(self._psi, self._nqubit, int(qubits[0]), int(qubits[1]))

    def permute(self, permutation: Sequence[int]) -> None:  # statevec.py:544-550
        n = self._nqubit
        if len(permutation) != n:
            raise ValueError(f"Permutation has length {len(permutation)}, but {n} qubits expected.")
        if set(permutation) != set(range(n)):
            raise ValueError(f"{permutation} is not a permutation.")
        t = np.transpose(self.psi.reshape((2,) * n), permutation)
        self._psi[: 1 << n] = t.reshape(1 << n)

    def fidelity(self, other) -> float:  # statevec.py:552-568
        return float(np.abs(np.dot(self.flatten().conjugate(), np.asarray(other.flatten()))) ** 2)

    def isclose(self, other, *, rtol: float = 1e-9, atol: float = 0.0) -> bool:
        return math.isclose(self.fidelity(other), 1, rel_tol=rtol, abs_tol=atol)


def _plus():
    from graphix_b200.mbqc import BasicStates  # noqa: PLC0415

    return BasicStates.PLUS


def _is_plus(data) -> bool:
    if not _is_state(data):
        return False
    v = np.asarray(data.to_statevector_numpy())
    return bool(np.array_equal(v, np.array([1, 1]) / np.sqrt(2)))


def _kron(states) -> np.ndarray:  # statevec.py:175-185
    psi = np.array([1], dtype=np.complex128)
    for s in states:
        if not _is_state(s):
            raise TypeError("Data should be an homogeneous sequence of states.")
        psi = np.kron(psi, np.asarray(s.to_statevector_numpy(), dtype=np.complex128))
    return psi


_PAULI = (
    np.array([[0, 1], [1, 0]], dtype=np.complex128),
    np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    np.array([[1, 0], [0, -1]], dtype=np.complex128),
)


def outcome_to_operator_matrix(vec, outcome: int) -> np.ndarray:
    """1/2 (I + (-1)^r v.sigma) — base_backend.py:506-542."""
    sign = (-1) ** int(outcome)
    op = np.eye(2, dtype=np.complex128) / 2
    for t, p in zip(vec, _PAULI, strict=True):
        op = op + sign * float(t) * p / 2
    return op


class OracleBackend:
    """DenseStateBackend restated — base_backend.py:739-898."""

    def __init__(self, state: OracleStatevector | None = None, *, node_index=None, branch_selector=None) -> None:
        from graphix_b200.backend import NodeIndex  # noqa: PLC0415  (pure-Python index map, no device code)
        from graphix_b200.mbqc import RandomBranchSelector  # noqa: PLC0415

        self.state = OracleStatevector(nqubit=0) if state is None else state
        self.node_index = NodeIndex() if node_index is None else node_index
        self.branch_selector = RandomBranchSelector() if branch_selector is None else branch_selector

    @classmethod
    def with_capacity(cls, max_qubits: int, state=None, **kwargs) -> "OracleBackend":  # statevec.py:807-832
        st = OracleStatevector(nqubit=0, max_qubits=max_qubits) if state is None else OracleStatevector(
            state, max_qubits=max_qubits)
        return cls(st, **kwargs)

    @property
    def nqubit(self) -> int:
        return self.state.nqubit

    def add_nodes(self, nodes, data=None) -> None:  # base_backend.py:778-797
        self.state.add_nodes(nqubit=len(nodes), data=_plus() if data is None else data)
        self.node_index.extend(nodes)

    def entangle_nodes(self, edge) -> None:  # base_backend.py:799-810
        self.state.entangle((self.node_index.index(edge[0]), self.node_index.index(edge[1])))

    def measure(self, node, measurement, rng=None, *, stacklevel: int = 1):  # base_backend.py:812-856
        loc = self.node_index.index(node)
        bloch = measurement.to_bloch()
        vec = bloch.plane.polar(bloch.angle)

        def f_expectation0() -> float:
            v = self.state.expectation_single(outcome_to_operator_matrix(vec, 0), loc)
            assert math.isclose(v.imag, 0, abs_tol=1e-10)
            return v.real

        outcome = self.branch_selector.measure(node, f_expectation0, rng, stacklevel=stacklevel + 1)
        self.state.project_qubit(outcome_to_operator_matrix(vec, outcome), loc)
        self.node_index.remove(node)
        return outcome

    def correct_byproduct(self, cmd) -> None:  # base_backend.py:858-862
        op = _PAULI[0] if cmd.kind.name == "X" else _PAULI[2]
        self.state.evolve_single(op, self.node_index.index(cmd.node))

    def apply_clifford(self, node, clifford) -> None:  # base_backend.py:881-885
        self.state.evolve_single(clifford.matrix, self.node_index.index(node))

    def apply_noise(self, cmd) -> None:
        raise RuntimeError("This backend does not support noise.")

    def finalize(self, output_nodes) -> None:  # base_backend.py:887-892
        self.state.permute([self.node_index.index(n) for n in output_nodes])
        self.node_index.clear()
        self.node_index.extend(output_nodes)
