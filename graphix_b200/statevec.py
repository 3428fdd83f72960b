"""``B200Statevector`` — the device-resident state object of the backend.

Mirrors the public surface of the reference ``Statevector``
(graphix/sim/statevec.py:52-750): same method names, argument meaning and
exceptions, so that ``DenseStateBackend``-style code and the reference's own
tests read the same.  Every numerical method is one call through the C ABI of
``include/gxsv.h`` into the sm_100a kernels; nothing here computes amplitudes
on the host.
"""
from __future__ import annotations

import ctypes as C
import math
from collections.abc import Iterable, Sequence
from numbers import Number

import numpy as np

from . import _lib
from .mbqc import BasicStates

_PLUS_VEC = np.array([1, 1], dtype=np.complex128) / math.sqrt(2)


def _noise_error_class():
    """graphix's own ``NoiseNotSupportedError`` when graphix is importable (so that ``except`` clauses written against
    graphix keep working), else an equivalent local class — graphix/sim/base_backend.py:324-329."""
    try:
        from graphix.sim.base_backend import NoiseNotSupportedError as Ref  # noqa: PLC0415

        return Ref
    except Exception:  # noqa: BLE001 - graphix absent
        class NoiseNotSupportedError(Exception):
            def __str__(self) -> str:
                return "This backend does not support noise."

        return NoiseNotSupportedError


NoiseNotSupportedError = _noise_error_class()


def _torch_stream_and_device(device: int | None) -> tuple[int, int | None]:
    """Current torch CUDA stream handle (0 when torch is not imported) and device index."""
    import sys  # noqa: PLC0415

    torch = sys.modules.get("torch")
    if torch is None or not torch.cuda.is_available():
        return 0, device
    if device is None:
        device = torch.cuda.current_device()
    return int(torch.cuda.current_stream(device).cuda_stream), device


def _is_state(obj: object) -> bool:
    return hasattr(obj, "to_statevector_numpy")


def _op8(op) -> C.Array:
    a = np.asarray(op)
    if a.dtype == np.object_:
        raise TypeError(
            "Statevector backend does not support symbolic operators. Consider using backend in `graphix-symbolic` plugin."
        )
    a = np.ascontiguousarray(a, dtype=np.complex128)
    if a.shape != (2, 2):
        raise ValueError(f"expected a 2x2 operator, got shape {a.shape}")
    return (C.c_double * 8)(*a.view(np.float64).ravel())


def ket_string(amplitudes: dict, precision: int = 4) -> str:
    """Plain ``a|00> + b|01>`` rendering of a ``to_dict()`` result (fallback of ``B200Statevector.draw``)."""
    if not amplitudes:
        return "0"
    terms = []
    for ket, a in amplitudes.items():
        a = complex(a)
        if abs(a.imag) < 10.0 ** -precision * max(abs(a.real), 1e-300):
            coeff = f"{a.real:.{precision}g}"
        elif abs(a.real) < 10.0 ** -precision * abs(a.imag):
            coeff = f"{a.imag:.{precision}g}i"
        else:
            coeff = f"({a.real:.{precision}g}{a.imag:+.{precision}g}i)"
        terms.append(f"{coeff}|{ket}>")
    return " + ".join(terms).replace("+ -", "- ")


class B200Statevector:
    """complex128 statevector in B200 HBM.

    Parameters follow ``Statevector.__init__`` (statevec.py:93-216): ``data`` is a
    ``State``, an iterable of ``State``, a ``2**n`` vector or another statevector;
    ``max_qubits`` preallocates capacity (``Pattern.max_space()``).
    Extra keyword-only parameters: ``device`` (CUDA ordinal), ``fusion``
    (0 = one kernel per command, 1 = runs of E fused, 2 = lazy N/E->M fusion [default]),
    ``strict`` (default True: the zero-norm ``RuntimeError`` is raised by the projection itself, like the reference;
    with ``strict=False`` it surfaces at ``sync()`` / ``flatten()`` / ``finalize``), ``batch`` (default 12 with
    ``fusion=2``: up to K consecutive fused projections run as ONE pass over the register — with ``strict=True`` only
    those whose outcome probability is known analytically, i.e. that cannot be a zero-norm branch; everything else
    still waits for its norm).
    """

    def __init__(self, data=BasicStates.PLUS, nqubit: int | None = None, max_qubits: int | None = None, *,
                 device: int | None = None, fusion: int = 2, strict: bool = True, batch: int | None = None) -> None:
        if nqubit is not None and nqubit < 0:
            raise ValueError("`nqubit` must be a non-negative integer.")
        if max_qubits is not None and max_qubits < 0:
            raise ValueError("`max_qubits` must be a non-negative integer.")
        vec, n = self._parse_data(data, nqubit)
        if max_qubits is not None and max_qubits < n:
            raise ValueError(f"`max_qubits` can't be smaller than the length of input state: {max_qubits} < {n}.")
        cap = n if max_qubits is None else max_qubits
        self._lib = _lib.load()
        stream, dev = _torch_stream_and_device(device)
        self._device = 0 if dev is None else dev
        self._stream = stream      # the CUDA stream every kernel of this state is launched on (0 = legacy default)
        self._h = C.c_void_p()
        rc = self._lib.gxsv_create(cap, self._device, C.c_void_p(stream), C.byref(self._h))
        if rc == _lib.GXSV_ENOMEM and fusion == 2 and cap > n and cap >= 30:
            # Lazy fusion materialises at most max_space - 1 qubits for the (N, E, M)* streams the transpiler emits
            # (the freshly prepared qubit takes over the measured qubit's amplitudes), so a register of half the
            # requested size is enough unless the pattern really needs the last bit — in which case the growth
            # fails later with the same MemoryError.  This is what lets a 34-live-qubit pattern run on one B200.
            import warnings  # noqa: PLC0415

            warnings.warn(f"2^{cap} amplitudes do not fit in device memory; allocating 2^{cap - 1} (lazy fusion "
                          "rarely needs the last qubit)", stacklevel=2)
            self._h = C.c_void_p()
            rc = self._lib.gxsv_create(cap - 1, self._device, C.c_void_p(stream), C.byref(self._h))
        if rc != _lib.GXSV_OK:
            self._h = C.c_void_p()
            _lib.raise_for(self._lib, None, rc)
        self._check(self._lib.gxsv_set_option(self._h, _lib.OPT_FUSION, int(fusion)))
        self._check(self._lib.gxsv_set_option(self._h, _lib.OPT_STRICT, int(bool(strict))))
        if batch is None:
            batch = 12 if fusion == 2 else 0
        if batch:
            if fusion != 2:
                raise ValueError("`batch` needs fusion=2")
            self._check(self._lib.gxsv_set_option(self._h, _lib.OPT_BATCH, int(batch)))
        if vec is not None:
            self._add(vec, n)

    # -- construction helpers ------------------------------------------------
    @staticmethod
    def _parse_data(data, nqubit: int | None):
        """Return (what to add, number of qubits). Same acceptance rules as statevec.py:125-205."""
        if isinstance(data, B200Statevector) or (hasattr(data, "flatten") and hasattr(data, "nqubit")
                                                 and not isinstance(data, np.ndarray)):
            flat = np.asarray(data.flatten(), dtype=np.complex128)
            if nqubit is not None and len(flat) != 1 << nqubit:
                raise ValueError(
                    f"Inconsistent parameters between nqubit = {nqubit} and the inferred number of qubit = {len(flat)}."
                )
            return flat, int(data.nqubit)
        if _is_state(data):
            n = 1 if nqubit is None else nqubit
            return [data] * n, n
        if isinstance(data, np.ndarray) and data.ndim == 1 and data.dtype.kind in "fc" and data.size > 0:
            # numeric vector fast path (no Python list of 2^n scalars); same checks as statevec.py:188-197
            length = int(data.size)
            inferred = length.bit_length() - 1
            if nqubit is None:
                if length & (length - 1):
                    raise ValueError(f"Length of input data is not a power of two: {length}.")
            elif nqubit != inferred:
                raise ValueError(f"Mismatch between nqubit and inferred nqubit: {nqubit} != {inferred}.")
            psi = np.ascontiguousarray(data, dtype=np.complex128)
            if not np.isclose(math.sqrt(np.vdot(psi, psi).real), 1.0):
                raise ValueError("Input state is not normalized.")
            return psi, inferred
        if not isinstance(data, Iterable):
            raise TypeError(f"Incorrect type for data: {type(data)}.")
        items = list(data)
        if len(items) == 0:
            if nqubit is not None and nqubit != 0:
                raise ValueError("`nqubit` is not null but input state is empty.")
            return None, 0
        if _is_state(items[0]):
            if nqubit is not None and nqubit != len(items):
                raise ValueError(f"Mismatch between nqubit and length of input state: {nqubit} != {len(items)}.")
            for s in items:
                if not _is_state(s):
                    raise TypeError("Data should be an homogeneous sequence of states.")
            return items, len(items)
        if isinstance(items[0], (Number, np.generic)):
            length = len(items)
            inferred = length.bit_length() - 1
            if nqubit is None:
                if length & (length - 1):
                    raise ValueError(f"Length of input data is not a power of two: {length}.")
            elif nqubit != inferred:
                raise ValueError(f"Mismatch between nqubit and inferred nqubit: {nqubit} != {inferred}.")
            psi = np.array(items, dtype=np.complex128)
            if not np.isclose(np.linalg.norm(psi), 1.0):
                raise ValueError("Input state is not normalized.")
            return psi, inferred
        raise TypeError(f"First element of data has type {type(items[0])} whereas Number or State is expected.")

    def _add(self, what, n: int) -> None:
        if isinstance(what, np.ndarray):
            buf = np.ascontiguousarray(what, dtype=np.complex128)
            self._check(self._lib.gxsv_tensor(self._h, n, buf.ctypes.data_as(C.c_void_p)))
            return
        for s in what:  # product state: one device-side append per qubit, no 2^n host vector
            if s is BasicStates.PLUS:
                self._check(self._lib.gxsv_add_plus(self._h))
                continue
            v = np.asarray(s.to_statevector_numpy(), dtype=np.complex128)
            if np.array_equal(v, _PLUS_VEC):
                self._check(self._lib.gxsv_add_plus(self._h))
            else:
                self._check(self._lib.gxsv_add_qubit(self._h, (C.c_double * 4)(v[0].real, v[0].imag, v[1].real, v[1].imag)))

    def _check(self, rc: int) -> None:
        if rc != _lib.GXSV_OK:
            _lib.raise_for(self._lib, self._h, rc)

    def __del__(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gxsv_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._h = C.c_void_p()

    # -- reference surface ---------------------------------------------------
    @property
    def nqubit(self) -> int:
        """statevec.py:232-236."""
        return self._lib.gxsv_nqubit(self._h)

    @property
    def max_qubits(self) -> int:
        """statevec.py:238-241."""
        return self._lib.gxsv_max_qubits(self._h)

    def ensure_capacity(self, required_qubits: int) -> None:
        """statevec.py:243-259."""
        self._check(self._lib.gxsv_ensure_capacity(self._h, int(required_qubits)))

    def flatten(self, out: np.ndarray | None = None) -> np.ndarray:
        """2**nqubit complex128 on the host, qubit 0 = MSB — statevec.py:261-267.

        ``out`` (optional) is a preallocated C-contiguous complex128 array, e.g. pinned memory."""
        size = 1 << self.nqubit
        if out is None:
            out = np.empty(size, dtype=np.complex128)
        elif out.dtype != np.complex128 or out.size != size or not out.flags.c_contiguous:
            raise ValueError(f"`out` must be a C-contiguous complex128 array of {size} elements")
        self._check(self._lib.gxsv_flatten(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def psi(self) -> np.ndarray:
        """statevec.py:222-229 (a host copy here, not a view)."""
        return self.flatten()

    @property
    def _psi(self) -> np.ndarray:
        """Host buffer of ``2**max_qubits`` amplitudes whose first ``2**nqubit`` entries are the state — what the
        reference's own copy constructor reads from a ``Statevector`` instance (``Statevector(other)``,
        statevec.py:128-135, used e.g. by ``DensityMatrix(statevec)``); this class is registered as a virtual
        ``Statevector`` subclass, so graphix code may take that path.  A copy, made on demand."""
        buf = np.empty(1 << self.max_qubits, dtype=np.complex128)
        self.flatten(out=buf[: 1 << self.nqubit])
        return buf

    def add_nodes(self, nqubit: int, data) -> None:
        """statevec.py:269-300."""
        if nqubit == 1 and data is BasicStates.PLUS:
            self._check(self._lib.gxsv_add_plus(self._h))
            return
        what, n = self._parse_data(data, nqubit)
        if what is not None:
            self._add(what, n)

    def entangle(self, qubits: tuple[int, int]) -> None:
        """CZ — statevec.py:302-315."""
        self._check(self._lib.gxsv_entangle(self._h, int(qubits[0]), int(qubits[1])))

    def evolve_single(self, op, qubit: int) -> None:
        """statevec.py:317-332."""
        self._check(self._lib.gxsv_evolve_single(self._h, int(qubit), _op8(op)))

    def expectation_single(self, op, qubit: int) -> complex:
        """statevec.py:334-359."""
        out = (C.c_double * 2)()
        self._check(self._lib.gxsv_expectation_single(self._h, int(qubit), _op8(op), out))
        return complex(out[0], out[1])

    def evolve(self, op, qubits: Sequence[int]) -> None:
        """statevec.py:361-395 (k-qubit operator; not on the pattern path)."""
        k = len(qubits)
        a = np.ascontiguousarray(np.asarray(op), dtype=np.complex128)
        if k == 1:
            self.evolve_single(a.reshape(2, 2), qubits[0])
            return
        q = (C.c_int * k)(*[int(x) for x in qubits])
        self._check(self._lib.gxsv_evolve(self._h, k, q, a.ctypes.data_as(C.c_void_p)))

    def expectation_value(self, op, qubits: Sequence[int]) -> complex:
        """<psi| op |psi> for a k-qubit operator — statevec.py:397-415 (same recipe: evolve a copy, then <psi|.>), all on
        the device: the copy is a second register (``gxsv_clone``), the inner product reads both registers in place."""
        other = self._clone()
        other.evolve(op, qubits)
        out = (C.c_double * 2)()
        self._check(self._lib.gxsv_inner(self._h, other._h, out))
        return complex(out[0], out[1])

    def _clone(self) -> "B200Statevector":
        new = object.__new__(B200Statevector)
        new._lib, new._device, new._stream = self._lib, self._device, self._stream
        new._h = C.c_void_p()
        self._check(self._lib.gxsv_clone(self._h, C.byref(new._h)))
        return new

    def remove_qubit(self, qubit: int) -> None:
        """statevec.py:417-428."""
        raise NotImplementedError("This method is deprecated. See :meth:`self.project_qubit`.")

    def project_qubit(self, op, qubit: int) -> None:
        """Project + remove + renormalise — statevec.py:430-495 (atol = 1e-10, :495)."""
        self._check(self._lib.gxsv_project_qubit(self._h, int(qubit), _op8(op), 1e-10, None))

    def swap(self, qubits: tuple[int, int]) -> None:
        """statevec.py:497-506."""
        self._check(self._lib.gxsv_swap(self._h, int(qubits[0]), int(qubits[1])))

    def tensor(self, other) -> None:
        """self (x) other — statevec.py:508-524."""
        flat = np.ascontiguousarray(other.flatten(), dtype=np.complex128)
        self._check(self._lib.gxsv_tensor(self._h, int(other.nqubit), flat.ctypes.data_as(C.c_void_p)))

    def permute(self, permutation: Sequence[int]) -> None:
        """statevec.py:544-550."""
        n = len(permutation)
        if n != self.nqubit:
            raise ValueError(f"Permutation has length {n}, but {self.nqubit} qubits expected.")
        if set(permutation) != set(range(n)):
            raise ValueError(f"{permutation} is not a permutation.")
        self._check(self._lib.gxsv_permute(self._h, (C.c_int * n)(*[int(p) for p in permutation]), n))

    def apply_noise(self, qubits: Sequence[int], noise) -> None:  # noqa: ARG002
        """base_backend.py:487-503: statevectors do not support noise."""
        raise NoiseNotSupportedError

    def fidelity(self, other) -> float:
        """|<self|other>|^2 — statevec.py:552-568; on device when both states are."""
        if isinstance(other, B200Statevector) and other._device == self._device:
            out = C.c_double()
            self._check(self._lib.gxsv_fidelity(self._h, other._h, C.byref(out)))
            return float(out.value)
        inner = np.dot(self.flatten().conjugate(), np.asarray(other.flatten()))
        return float(np.abs(inner) ** 2)

    def isclose(self, other, *, rtol: float = 1e-09, atol: float = 0.0) -> bool:
        """statevec.py:570-589."""
        return math.isclose(self.fidelity(other), 1, rel_tol=rtol, abs_tol=atol)

    def to_dict(self, encoding: str = "MSB", *, rtol: float = 0.0, atol: float = 1e-8) -> dict:
        """statevec.py:590-640."""
        return self._to_dict_map(lambda x: x, encoding, rtol=rtol, atol=atol)

    def to_prob_dict(self, encoding: str = "MSB", *, rtol: float = 0.0, atol: float = 1e-8) -> dict:
        """statevec.py:642-668."""
        return self._to_dict_map(lambda x: np.abs(x) ** 2, encoding, rtol=rtol, atol=atol)

    def draw(self, output=None, *, encoding: str = "MSB", max_denominator: int = 1000, atol: float = 1e-9,
             rtol: float = 0.0, precision: int = 4) -> str:
        """Ket-notation string — statevec.py:680-739.

        Formatting is host-side pretty printing, not part of the device path: with graphix installed its own
        ``graphix.pretty_print.statevec_to_str`` is used (it only reads ``to_dict`` from the state); without it the
        amplitudes are printed as decimals.
        """
        try:
            from graphix.pretty_print import statevec_to_str  # noqa: PLC0415
        except ImportError:
            return ket_string(self.to_dict(encoding, rtol=rtol, atol=atol), precision)
        return statevec_to_str(self, output, encoding=encoding, max_denominator=max_denominator, atol=atol, rtol=rtol,
                               precision=precision)

    SPARSE_CAP = 1 << 20     # entries a sparse read-out may return before falling back to a full flatten()

    def _to_dict_map(self, f, encoding: str, *, rtol: float, atol: float) -> dict:
        """statevec.py:591-678.  The entries with ``|amp| > atol`` (``np.isclose(|amp|, 0, rtol, atol)`` compares with
        ``atol + rtol * 0``) are compacted on the device — only they cross to the host, not the 2**n amplitudes."""
        del rtol
        n = self.nqubit
        lib = getattr(self, "_lib", None)
        keep = vals = None
        if lib is not None and hasattr(lib, "gxsv_nonzero"):
            cap = min(1 << n, self.SPARSE_CAP)
            idx = np.empty(cap, dtype=np.uint64)
            amp = np.empty(cap, dtype=np.complex128)
            count = C.c_uint64()
            self._check(lib.gxsv_nonzero(self._h, float(atol), cap, idx.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         amp.ctypes.data_as(C.POINTER(C.c_double)), C.byref(count)))
            if count.value <= cap:
                order = np.argsort(idx[: count.value], kind="stable")
                keep, vals = idx[: count.value][order].astype(np.int64), f(amp[: count.value][order])
        if keep is None:
            flat = self.flatten()
            keep = np.flatnonzero(np.abs(flat) > atol)
            vals = f(flat[keep])
        keys = [format(int(i), f"0{n}b") if n else "0" for i in keep]
        if encoding == "LSB":
            keys = [k[::-1] for k in keys]
        return dict(zip(keys, vals, strict=True))

    def __str__(self) -> str:
        return f"B200Statevector(nqubit={self.nqubit}, max_qubits={self.max_qubits}, device={self._device})"

    __repr__ = __str__

    # -- engine plumbing -------------------------------------------------------
    def norm2(self) -> float:
        out = C.c_double()
        self._check(self._lib.gxsv_norm2(self._h, C.byref(out)))
        return float(out.value)

    def flush(self) -> None:
        self._check(self._lib.gxsv_flush(self._h))

    def sync(self) -> None:
        self._check(self._lib.gxsv_sync(self._h))

    def reset(self) -> None:
        self._check(self._lib.gxsv_reset(self._h))

    def set_option(self, key: int, value: int) -> None:
        self._check(self._lib.gxsv_set_option(self._h, key, value))

    def layout(self) -> list[int]:
        n = self.nqubit
        out = (C.c_int * max(n, 1))()
        self._check(self._lib.gxsv_layout(self._h, out))
        return list(out[:n])

    def stats(self) -> dict[str, dict[str, float]]:
        st = _lib.Stats()
        self._check(self._lib.gxsv_stats(self._h, C.byref(st)))
        return {
            name: {"launches": int(st.launches[i]), "algo_bytes": float(st.algo_bytes[i]),
                   "device_ms": float(st.device_ms[i])}
            for i, name in enumerate(_lib.KIND_NAMES)
        }

    def reset_stats(self) -> None:
        self._check(self._lib.gxsv_reset_stats(self._h))


def _register_with_graphix() -> None:
    """When graphix is importable, make ``isinstance(state, graphix.sim.statevec.Statevector)`` (and ``DenseState``)
    hold for ``B200Statevector`` — helper code in graphix and its tests dispatches on those classes
    (e.g. tests/test_pattern.py:39-44).  ``Statevector`` is an ABC (via ``DenseState``), so this is a virtual
    subclass registration: no graphix code is inherited."""
    try:
        from graphix.sim.base_backend import DenseState  # noqa: PLC0415
        from graphix.sim.statevec import Statevector  # noqa: PLC0415
    except Exception:  # noqa: BLE001 - graphix absent
        return
    DenseState.register(B200Statevector)
    Statevector.register(B200Statevector)


_register_with_graphix()
