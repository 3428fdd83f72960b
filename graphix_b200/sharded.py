"""Statevector sharded over ranks by its top (rank-index) qubits — SURVEY.md §8(e).

One process per GPU (``torchrun``); every rank walks the same command stream and takes the same
branch.  Rank ``r`` holds the sub-cube of the amplitude index whose *global* bits equal the bits
of ``r``; the remaining bits are the local engine's (``libgxsv`` in ``GXSV_OPT_DIST`` mode, lazy
fusion on).  This module is host logic only — layout decisions, the exchange schedule and the
collectives — and is engine-agnostic, so the same code runs under ``gloo`` on CPU with the
test double of ``tests/cpu_shard_engine.py``.

Rules
* A global bit with no qubit on it is *replicated*: the ranks differing only in that bit hold the
  same shard and do the same work, so putting a freshly prepared qubit on it costs **no
  communication** (each rank scales by its own amplitude).  Qubits go to the local engine while it
  has room (``local_bits``) and to global bits after that.
* Commands on local qubits run on the shard with no communication.  CZ never communicates: between
  two global qubits it is a per-rank sign, between a local and a global qubit it is a Z on the
  ranks whose bit is 1 — queued and folded into the next projection of that local qubit (or turned
  back into a local CZ if the global qubit is swapped into the shard first).
* Z / diagonal gates on a global qubit are per-rank scalars, X is a relabel of the rank bit.
* Probabilities and norms: one small all-reduce.
* A projection (or any non-diagonal 2x2) on a global qubit first swaps it with a local qubit: each
  rank sends one half of its shard to the partner rank and receives the other half (NCCL
  send/recv over NVLink, chunked through a bounded staging buffer).  Global bits therefore never
  die, and every rank keeps an equal share of the state when qubits are removed.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections.abc import Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .backend import B200StatevectorBackend, NodeIndex
from .mbqc import BasicStates, RandomBranchSelector
from .statevec import B200Statevector, NoiseNotSupportedError, _op8

_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)


class GpuShardEngine:
    """The local shard: a ``B200Statevector`` in distributed mode plus torch staging buffers."""

    def __init__(self, local_bits: int, device: int | None = None, strict: bool = True, batch: int = 12) -> None:
        self.sv = B200Statevector(nqubit=0, max_qubits=local_bits, device=device, fusion=2, strict=strict,
                                  batch=0 if strict else batch)   # strict sharded projections wait for the all-reduced norm
        self.sv.set_option(_lib.OPT_DIST, 1)
        if strict and batch:
            # strict sharded mode: projections whose outcome probability is analytic (they cannot fail) are queued and
            # share passes; project_local then returns NaN norms and the orchestrator collects the norms later
            self.sv.set_option(_lib.OPT_BATCH, batch)
            self.sv.set_option(_lib.OPT_DIST_QUEUE, 1)
        # what was really allocated: the constructor may have fallen back to half the requested register (lazy fusion
        # rarely needs the last qubit); qubits beyond it go to free rank bits instead of failing later
        self.capacity = self.sv.max_qubits
        self.device = torch.device("cuda", self.sv._device)
        self._lib = self.sv._lib
        self._h = self.sv._h
        self.strict = strict

    # queries
    @property
    def nqubit(self) -> int:
        return self.sv.nqubit

    @property
    def nreal(self) -> int:
        return self._lib.gxsv_nreal(self._h)

    def is_virtual(self, q: int) -> bool:
        return self._lib.gxsv_is_virtual(self._h, q) == 1

    def half_size(self) -> int:
        return 1 << (self.nreal - 1)

    # commands
    def add_qubit(self, s0: complex, s1: complex) -> None:
        self.sv._check(self._lib.gxsv_add_qubit(self._h, (C.c_double * 4)(s0.real, s0.imag, s1.real, s1.imag)))

    def tensor(self, vec: np.ndarray) -> None:
        """Append log2(len(vec)) qubits holding ``vec`` (first new qubit = MSB; no normalisation check: a rank's
        slice of a joint input vector is not normalised on its own)."""
        buf = np.ascontiguousarray(vec, dtype=np.complex128)
        k = int(buf.size).bit_length() - 1
        self.sv._check(self._lib.gxsv_tensor(self._h, k, buf.ctypes.data_as(C.c_void_p)))

    def entangle(self, a: int, b: int) -> None:
        self.sv.entangle((a, b))

    def evolve_single(self, q: int, m) -> None:
        self.sv.evolve_single(m, q)

    def expectation_local(self, q: int, m) -> complex:
        return self.sv.expectation_single(m, q)

    def project_local(self, q: int, m) -> tuple[float, float]:
        norms = (C.c_double * 2)()
        self.sv._check(self._lib.gxsv_project_qubit(self._h, q, _op8(m), 0.0, norms))
        return norms[0], norms[1]

    def set_norm(self, total: float) -> None:
        self.sv._check(self._lib.gxsv_set_norm(self._h, total))

    def project_local_async(self, q: int, m) -> None:
        """Queue / launch the projection without waiting (consecutive fused projections share one pass)."""
        self.sv._check(self._lib.gxsv_project_qubit(self._h, q, _op8(m), 0.0, None))

    def take_norms(self) -> list[float]:
        """Flush, wait, and return the local norm^2 after every projection issued since the last call."""
        cap = 4096
        buf = (C.c_double * cap)()
        n = C.c_int()
        self.sv._check(self._lib.gxsv_take_norms(self._h, buf, cap, C.byref(n)))
        return list(buf[: n.value])

    def scale_host(self, c: complex) -> None:
        self.sv._check(self._lib.gxsv_scale_host(self._h, c.real, c.imag))

    def scale(self, c: complex) -> None:
        self.sv._check(self._lib.gxsv_scale(self._h, c.real, c.imag))

    def settle(self, q: int) -> None:
        self.sv._check(self._lib.gxsv_settle(self._h, q))

    def pack_half(self, q: int, bit: int, first: int, count: int, out: torch.Tensor) -> None:
        self.sv._check(self._lib.gxsv_pack_half(self._h, q, bit, first, count, C.c_void_p(out.data_ptr())))

    def unpack_half(self, q: int, bit: int, first: int, count: int, src: torch.Tensor, factor: complex) -> None:
        self.sv._check(self._lib.gxsv_unpack_half(self._h, q, bit, first, count, C.c_void_p(src.data_ptr()),
                                                  factor.real, factor.imag))

    # peer-memory exchange (CUDA IPC over NVLink): see include/gxsv.h
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.sv._check(self._lib.gxsv_ipc_export(self._h, buf))
        return buf.raw

    def ipc_open(self, slot: int, handle: bytes) -> None:
        self.sv._check(self._lib.gxsv_ipc_open(self._h, slot, handle))

    def ipc_close(self) -> None:
        self._lib.gxsv_ipc_close(self._h)

    def swap_peer(self, q: int, slot: int, my_bit: int, factor: complex) -> None:
        self.sv._check(self._lib.gxsv_swap_peer(self._h, q, slot, my_bit, factor.real, factor.imag))

    def stream_handle(self) -> int:
        return self.sv._stream

    def flush_batch(self) -> None:
        self.sv._check(self._lib.gxsv_flush_batch(self._h))

    def norm2_local(self) -> float:
        return self.sv.norm2()

    def flatten_local(self) -> np.ndarray:
        return self.sv.flatten()

    def reset(self) -> None:
        self.sv.reset()

    def sync(self) -> None:
        self.sv.sync()

    def stats(self):
        return self.sv.stats()

    def reset_stats(self) -> None:
        self.sv.reset_stats()

    def set_option(self, key: int, value: int) -> None:
        self.sv.set_option(key, value)


class ShardedInput:
    """A joint input state of ``nqubit`` qubits handed over shard by shard, for inputs too large (or too wasteful) to
    build in full on every rank: ``slice_for(t, kg, kl)`` returns the ``2**kl`` amplitudes (C-contiguous complex128,
    ideally pinned host memory) of the joint vector whose first ``kg`` qubits (first = MSB) have the value ``t``.  Accepted
    by ``ShardedStatevector.add_nodes`` / ``ShardedStatevectorBackend.add_nodes`` in place of a ``2**nqubit`` vector
    (graphix/sim/statevec.py:186-205); normalisation is the caller's responsibility."""

    def __init__(self, nqubit: int, slice_for) -> None:
        self.nqubit = int(nqubit)
        self.slice_for = slice_for


class _Q:
    """Location of one logical qubit: local engine index ``li`` or global bit ``g``."""

    __slots__ = ("li", "g", "z")

    def __init__(self, li: int | None = None, g: int | None = None) -> None:
        self.li = li
        self.g = g
        self.z = 0      # parity of rank-uniform Z's queued on a local qubit (X on a global qubit pushed through a CZ)


class ShardedStatevector:
    """State object with the reference ``Statevector`` surface over P ranks (P a power of two)."""

    CHUNK_MIN = 1 << 25   # amplitudes per staging buffer: at least 512 MiB (measured: 128 MiB chunks halve the exchange
    # rate — every torch.distributed P2P batch has a fixed cost of a few hundred microseconds on both sides) ...
    CHUNK_MAX = 1 << 27   # ... at most 2 GiB; two send + two receive buffers, allocated on demand
    WINDOW = 256     # non-strict mode: projections between two renormalisations (one small all-reduce each; the stored
    # amplitudes grow by at most a factor 2 per projection in between: 2^256 is far inside the double range)

    def __init__(self, engine, group=None, strict: bool | None = None) -> None:
        self.engine = engine
        self.group = group
        # strict: every projection waits for the all-reduced norm (zero-norm RuntimeError raised by the call
        # itself, like the reference).  Non-strict: nothing waits; norms are checked at the next sync point.
        self.strict = getattr(engine, "strict", True) if strict is None else strict
        self._window: list[tuple] = []   # per queued projection: (weight, big row, |row ratio|^2, phase, assumed, qubit)
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.P & (self.P - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.gbits = self.P.bit_length() - 1
        self._stage: dict[str, torch.Tensor] = {}
        self.exchanged_bytes = 0
        self.n_swaps = 0
        self.victim_priority = None     # optional callable(logical index) -> float: larger = evict to a global bit first
        self.profile = False            # bracket every exchange with CUDA events (bench.py)
        self._xev: list[tuple] = []
        self.exchange_ms = 0.0
        self.exchange_host_ms = 0.0
        self.sync_wait_ms = 0.0         # host time spent waiting for the device at deferred-check points
        self.transport = "nccl send/recv (pack -> staging -> unpack)"
        self._bar = None
        self._setup_peer_exchange()
        self._clear()

    def _setup_peer_exchange(self) -> None:
        """Map the partner ranks' registers (CUDA IPC) so that a qubit swap is ONE in-place kernel over NVLink on each
        side instead of pack -> ncclSend/Recv -> unpack through staging buffers.  All ranks use it or none does."""
        eng = self.engine
        if self.P == 1 or not hasattr(eng, "ipc_export") or os.environ.get("GXSV_EXCHANGE", "peer") != "peer":
            return
        ok = 1
        handles = None
        try:
            t = torch.frombuffer(bytearray(eng.ipc_export()), dtype=torch.uint8).to(eng.device)
            handles = [torch.empty_like(t) for _ in range(self.P)]
            dist.all_gather(handles, t, group=self.group)
        except Exception:  # noqa: BLE001 - any failure (IPC unsupported, ...) means: keep the NCCL path
            ok = 0
        if ok:
            try:
                for k in range(self.gbits):
                    eng.ipc_open(k, bytes(handles[self.rank ^ (1 << k)].cpu().numpy().tobytes()))
            except Exception:  # noqa: BLE001
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.transport = "peer load/store (CUDA IPC over NVLink), in place, no staging"
            self._bar = torch.zeros(1, dtype=torch.float32, device=eng.device)
        else:
            eng.ipc_close()

    def _stream_barrier(self) -> None:
        """Cross-rank barrier ordered on the compute stream (no host wait): a one-element all-reduce."""
        dist.all_reduce(self._bar, group=self.group)

    def _clear(self) -> None:
        self.q: list[_Q] = []                      # logical order (reference convention: index 0 = MSB)
        self.owner: list[_Q | None] = [None] * self.gbits   # qubit on each global bit (None = replicated)
        self.flip = [0] * self.gbits               # logical value of a global qubit = rank bit ^ flip
        self.u = np.ones(self.P, dtype=np.complex128)       # rank-dependent scalar of every rank's shard
        self.pend_lg: list[tuple[_Q, _Q]] = []     # queued CZ(local, global)

    # -- helpers -----------------------------------------------------------------------
    def _eb(self, k: int, r: int | None = None) -> int:
        r = self.rank if r is None else r
        return ((r >> k) & 1) ^ self.flip[k]

    def _dead(self) -> int:
        return sum(1 for o in self.owner if o is None)

    def _allreduce(self, vals: Sequence[float]) -> list[float]:
        if self.P == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.engine.device)
        dist.all_reduce(t, group=self.group)
        return t.tolist()

    def _sigma(self, qb: _Q) -> int:
        """Parity of the Z's queued on local qubit ``qb`` for this rank."""
        s = qb.z
        for a, b in self.pend_lg:
            if a is qb:
                s ^= self._eb(b.g)
        return s

    def _drop_pending(self, qb: _Q) -> None:
        qb.z = 0
        self.pend_lg = [(a, b) for (a, b) in self.pend_lg if a is not qb]

    def _flush_local(self, qb: _Q) -> None:
        """Apply the queued Z's of local qubit ``qb`` for real."""
        if any(a is qb for a, _ in self.pend_lg) and self.engine.is_virtual(qb.li):
            # The Z below differs between ranks.  A qubit the engine still holds as a product factor outside the
            # stored amplitudes must stay the same on every rank — the half-shard exchange moves stored amplitudes
            # only, and the engine keeps part of a factor in a host scalar that has to be rank-uniform — so the
            # qubit is materialised first (on every rank, whatever this rank's parity is).
            self.engine.settle(qb.li)
        if self._sigma(qb):
            self.engine.evolve_single(qb.li, _Z)
        self._drop_pending(qb)

    def _check(self, q: int) -> _Q:
        if not 0 <= q < len(self.q):
            raise IndexError(f"Qubit index {q} out of range [0, {len(self.q) - 1}]")
        return self.q[q]

    # -- reference surface ---------------------------------------------------------------
    @property
    def nqubit(self) -> int:
        return len(self.q)

    @property
    def max_qubits(self) -> int:
        return self.engine.capacity + self.gbits

    def add_nodes(self, nqubit: int, data) -> None:
        """statevec.py:269-300: product states qubit by qubit, joint 2^k vectors sliced over the ranks."""
        if isinstance(data, ShardedInput):
            if nqubit is not None and nqubit != data.nqubit:
                raise ValueError(f"Mismatch between nqubit and inferred nqubit: {nqubit} != {data.nqubit}.")
            self._add_vector(None, sharded=data)
            return
        vec = self._joint_vector(nqubit, data)
        if vec is not None:
            self._add_vector(vec)
            return
        states = self._states_of(nqubit, data)
        for s0, s1 in states:
            if self.engine.nqubit < self.engine.capacity:
                self.engine.add_qubit(s0, s1)
                self.q.append(_Q(li=self.engine.nqubit - 1))
                continue
            free = [k for k in range(self.gbits) if self.owner[k] is None]
            if not free:
                raise MemoryError(f"sharded state is full: {self.max_qubits} qubits")
            k = free[0]
            # replicated bit becomes live: every rank keeps its copy and takes its own amplitude
            qb = _Q(g=k)
            self.owner[k] = qb
            self.flip[k] = 0
            mine = s1 if (self.rank >> k) & 1 else s0
            if s0 == 0 or s1 == 0:
                self.engine.scale(complex(mine))      # exact zeros cannot live in the bookkeeping scalar
            else:
                for r in range(self.P):
                    self.u[r] *= s1 if (r >> k) & 1 else s0
            self.q.append(qb)

    @staticmethod
    def _joint_vector(nqubit: int, data) -> np.ndarray | None:
        """A 2^k amplitude vector (k >= 2) handed to ``add_nodes`` — statevec.py:186-205 — else None."""
        if hasattr(data, "to_statevector_numpy"):
            return None
        if hasattr(data, "flatten") and hasattr(data, "nqubit") and not isinstance(data, np.ndarray):
            vec = np.asarray(data.flatten(), dtype=np.complex128)       # a Statevector-like object
        else:
            if not isinstance(data, np.ndarray):
                data = list(data)
                if not data or hasattr(data[0], "to_statevector_numpy"):
                    return None
            vec = np.asarray(data, dtype=np.complex128).reshape(-1)
        if vec.size <= 2:
            return None
        if vec.size & (vec.size - 1):
            raise ValueError(f"Length of input data is not a power of two: {vec.size}.")
        inferred = int(vec.size).bit_length() - 1
        if nqubit is not None and nqubit != inferred:
            raise ValueError(f"Mismatch between nqubit and inferred nqubit: {nqubit} != {inferred}.")
        if not np.isclose(np.linalg.norm(vec), 1.0):
            raise ValueError("Input state is not normalized.")
        return vec

    def _add_vector(self, vec: np.ndarray | None, sharded: ShardedInput | None = None) -> None:
        """Tensor a joint vector of k new qubits onto the state.  The last qubits go to the shard while it has room,
        the first ones take free rank bits: every rank keeps the slice of ``vec`` that its rank bits select."""
        k = sharded.nqubit if sharded is not None else int(vec.size).bit_length() - 1
        kl = min(k, self.engine.capacity - self.engine.nqubit)
        kg = k - kl
        free = [b for b in range(self.gbits) if self.owner[b] is None]
        if kg > len(free):
            raise MemoryError(f"sharded state is full: {self.max_qubits} qubits")
        new_q = []
        t = 0
        for i in range(kg):
            b = free[i]
            qb = _Q(g=b)
            self.owner[b] = qb
            self.flip[b] = 0
            t = (t << 1) | ((self.rank >> b) & 1)
            new_q.append(qb)
        if sharded is not None:
            mine = np.asarray(sharded.slice_for(t, kg, kl))
            if mine.shape != (1 << kl,):
                raise ValueError(f"ShardedInput.slice_for returned shape {mine.shape}, expected ({1 << kl},)")
        else:
            mine = vec.reshape(1 << kg, 1 << kl)[t]
        first = self.engine.nqubit
        if kl:
            self.engine.tensor(mine)
            new_q.extend(_Q(li=first + j) for j in range(kl))
        else:
            self.engine.scale(complex(mine[0]))      # may be exactly zero: it cannot live in the bookkeeping scalar
        self.q.extend(new_q)

    @staticmethod
    def _states_of(nqubit: int, data) -> list[tuple[complex, complex]]:
        if hasattr(data, "to_statevector_numpy"):
            v = np.asarray(data.to_statevector_numpy(), dtype=np.complex128)
            return [(complex(v[0]), complex(v[1]))] * nqubit
        items = list(data)
        if items and hasattr(items[0], "to_statevector_numpy"):
            if len(items) != nqubit:
                raise ValueError(f"Mismatch between nqubit and length of input state: {nqubit} != {len(items)}.")
            out = []
            for s in items:
                v = np.asarray(s.to_statevector_numpy(), dtype=np.complex128)
                out.append((complex(v[0]), complex(v[1])))
            return out
        v = np.asarray(items, dtype=np.complex128)
        if v.size == 2 and nqubit == 1:
            if not np.isclose(np.linalg.norm(v), 1.0):
                raise ValueError("Input state is not normalized.")
            return [(complex(v[0]), complex(v[1]))]
        raise NotImplementedError("sharded states accept product states (State objects) only")

    def entangle(self, qubits: tuple[int, int]) -> None:
        a, b = self._check(qubits[0]), self._check(qubits[1])
        if a is b:
            self.evolve_single(_Z, qubits[0])      # statevec.py:930-937 with equal indices: Z on the qubit
            return
        if a.g is None and b.g is None:
            self.engine.entangle(a.li, b.li)
        elif a.g is not None and b.g is not None:
            for r in range(self.P):
                if self._eb(a.g, r) & self._eb(b.g, r):
                    self.u[r] = -self.u[r]
        else:
            loc, glo = (a, b) if a.g is None else (b, a)
            # Z on `loc` for the ranks whose bit of `glo` is 1.  Always queued: it is folded into the next
            # projection of `loc` for free, and it must follow `glo` if that qubit is swapped into the shard.
            self.pend_lg.append((loc, glo))

    def evolve_single(self, op, qubit: int) -> None:
        qb = self._check(qubit)
        m = np.asarray(op, dtype=np.complex128)
        diagonal = m[0, 1] == 0 and m[1, 0] == 0
        if qb.g is not None:
            k = qb.g
            if diagonal:
                for r in range(self.P):
                    self.u[r] *= m[1, 1] if self._eb(k, r) else m[0, 0]
                return
            if m[0, 0] == 0 and m[1, 1] == 0:
                # [[0, a], [b, 0]] = X . diag(b, a): per-rank scalar, then relabel the rank bit.
                for r in range(self.P):
                    self.u[r] *= m[0, 1] if self._eb(k, r) else m[1, 0]
                # X_k CZ_{lk} = CZ_{lk} Z_l X_k for every CZ still queued on k
                for a, b in self.pend_lg:
                    if b is qb:
                        a.z ^= 1
                self.flip[k] ^= 1
                return
            if self.engine.nqubit == 0:
                self._renorm()
                self._evolve_global_scalar(qb, m)
                return
            self._swap_in(qb)
        if diagonal:
            self.engine.evolve_single(qb.li, m)
            return
        self._flush_local(qb)
        self.engine.evolve_single(qb.li, m)

    def expectation_single(self, op, qubit: int) -> complex:
        qb = self._check(qubit)
        self._renorm()
        if qb.g is not None:
            if self.engine.nqubit == 0:
                return self._expect_global_scalar(qb, np.asarray(op, dtype=np.complex128))
            self._swap_in(qb)
        m = np.array(op, dtype=np.complex128)
        if self._sigma(qb):
            m[0, 1], m[1, 0] = -m[0, 1], -m[1, 0]     # Z op Z
        loc = self.engine.expectation_local(qb.li, m)
        w = abs(self.u[self.rank]) ** 2
        re, im = self._allreduce([w * loc.real, w * loc.imag])
        scale = 1.0 / (1 << self._dead())
        return complex(re * scale, im * scale)

    def project_qubit(self, op, qubit: int) -> None:
        """statevec.py:430-495 for rank-1 ``op`` (the projectors ``measure`` builds)."""
        qb = self._check(qubit)
        if qb.g is not None:
            if self.engine.nqubit == 0:
                self._renorm()
                self._project_global_scalar(qb, np.asarray(op, dtype=np.complex128))
                return
            self._swap_in(qb)
        m = np.array(op, dtype=np.complex128)
        if self._sigma(qb):
            m[0, 1], m[1, 1] = -m[0, 1], -m[1, 1]     # op . Z: the queued Z's act first
        self._drop_pending(qb)
        r0 = abs(m[0, 0]) ** 2 + abs(m[0, 1]) ** 2
        r1 = abs(m[1, 0]) ** 2 + abs(m[1, 1]) ** 2
        big = 0 if r0 >= r1 else 1
        w = abs(self.u[self.rank]) ** 2
        scale = 1.0 / (1 << self._dead())
        li = qb.li
        # phase convention: the reference keeps half 0 of the projected pair unless its norm is <= 1e-10
        # (statevec.py:1187-1190).  For a queued projection: guess now, correct in _renorm() once the norm is known.
        phi, assumed = 1.0 + 0j, False
        if big == 1:
            c = 0 if abs(m[1, 0]) >= abs(m[1, 1]) else 1
            ratio = m[0, c] / m[1, c]
            phi = complex(ratio / abs(ratio)) if abs(ratio) > 0 else 1.0 + 0j
            assumed = bool(r0 > 1e-10 * r1)
        if not self.strict:
            # non-blocking: the engine queues / launches the projection (consecutive ones share a pass) and nothing
            # waits; the norms of a whole window are all-reduced, checked and used to renormalise in _renorm()
            self.engine.project_local_async(li, m)
            n0l = n1l = math.nan
        else:
            # strict: the engine returns the local half-norms — or NaN when it queued the projection because its outcome
            # probability is analytic (it cannot be the zero-norm branch, nothing has to be waited for)
            n0l, n1l = self.engine.project_local(li, m)
        self._remove(qb)
        for o in self.q:
            if o.li is not None and o.li > li:
                o.li -= 1
        if math.isnan(n0l):
            if big == 1 and assumed:
                self.engine.scale_host(phi)
            self._window.append((w * scale, big, (r0 / r1) if big else (r1 / r0 if r0 > 0 else 0.0), phi, assumed, qubit))
            if len(self._window) >= self.WINDOW:
                self._renorm()
            return
        # synchronous projection; queued projections still pending are resolved in the same all-reduce
        prev, (n0, n1) = self._renorm(extra=[w * n0l * scale, w * n1l * scale], set_norm=False)
        pick = 0
        if n0 <= 1e-10 * prev:
            pick = 1
            if n1 <= 1e-10 * prev:
                raise RuntimeError(f"Attempted to remove qubit {qubit} from 0-norm statevector.")
        self.engine.set_norm(n1 if big else n0)
        if pick != big:
            row_b, row_p = m[big], m[pick]
            c = 0 if abs(row_b[0]) >= abs(row_b[1]) else 1
            ratio = row_p[c] / row_b[c]
            if abs(ratio) > 0:
                self.engine.scale_host(complex(ratio / abs(ratio)))

    def _remove(self, qb: _Q) -> None:
        self.q = [o for o in self.q if o is not qb]

    def _global_scalars(self) -> np.ndarray:
        """Every remaining qubit is global: the shards are scalars; gather them (true amplitudes)."""
        val = complex(self.engine.flatten_local()[0]) * self.u[self.rank]
        re = self._allgather_scalar(val.real)
        im = self._allgather_scalar(val.imag)
        return np.array(re) + 1j * np.array(im)

    def _evolve_global_scalar(self, qb: _Q, m: np.ndarray) -> None:
        k = qb.g
        vals = self._global_scalars()
        out = np.zeros(self.P, dtype=np.complex128)
        for r in range(self.P):
            e = self._eb(k, r)
            partner = r ^ (1 << k)
            v = [0j, 0j]
            v[e], v[1 - e] = vals[r], vals[partner]
            out[r] = m[e, 0] * v[0] + m[e, 1] * v[1]
        self.engine.reset()
        self.u = out

    def _expect_global_scalar(self, qb: _Q, m: np.ndarray) -> complex:
        k = qb.g
        vals = self._global_scalars()
        tot = 0j
        for r in range(self.P):
            if (r >> k) & 1:
                continue
            lo, hi = r, r | (1 << k)
            v = (vals[lo], vals[hi]) if self.flip[k] == 0 else (vals[hi], vals[lo])
            for a in (0, 1):
                for b in (0, 1):
                    tot += np.conj(v[a]) * m[a, b] * v[b]
        return complex(tot / (1 << self._dead()))

    def _project_global_scalar(self, qb: _Q, m: np.ndarray) -> None:
        """Edge case: every remaining qubit is global, the shards are scalars — contract on the host."""
        k = qb.g
        vals = self._global_scalars()
        r0 = abs(m[0, 0]) ** 2 + abs(m[0, 1]) ** 2
        r1 = abs(m[1, 0]) ** 2 + abs(m[1, 1]) ** 2
        big = 0 if r0 >= r1 else 1
        out = np.zeros(self.P, dtype=np.complex128)
        for r in range(self.P):
            lo = r & ~(1 << k)
            hi = r | (1 << k)
            v0, v1 = (vals[lo], vals[hi]) if self.flip[k] == 0 else (vals[hi], vals[lo])
            out[r] = m[big, 0] * v0 + m[big, 1] * v1
        self.owner[k] = None
        self.flip[k] = 0
        self._remove(qb)
        nbig = float(np.sum(np.abs(out) ** 2)) / (1 << self._dead())
        n0, n1 = (nbig, nbig * r1 / r0) if big == 0 else (nbig * r0 / r1, nbig)
        if n0 <= 1e-10 and n1 <= 1e-10:
            raise RuntimeError("Attempted to remove qubit from 0-norm statevector.")
        pick = 0 if n0 > 1e-10 else 1
        phase = 1.0
        if pick != big:
            c = 0 if abs(m[big, 0]) >= abs(m[big, 1]) else 1
            ratio = m[pick, c] / m[big, c]
            phase = ratio / abs(ratio) if abs(ratio) > 0 else 1.0
        self.engine.reset()
        self.u = out * phase / math.sqrt(nbig)
        self.u[self.u == 0] = 0

    def _allgather_scalar(self, x: float) -> list[float]:
        t = torch.zeros(self.P, dtype=torch.float64, device=self.engine.device)
        t[self.rank] = x
        if self.P > 1:
            dist.all_reduce(t, group=self.group)
        return t.tolist()

    # -- the exchange ------------------------------------------------------------------------
    def _staging(self, name: str, count: int) -> torch.Tensor:
        """``count`` amplitudes as 2*count float64 (NCCL has no complex types: (re, im) pairs travel as doubles)."""
        t = self._stage.get(name)
        if t is None or t.numel() < 2 * count:
            t = torch.empty(max(2 * count, 2), dtype=torch.float64, device=self.engine.device)
            self._stage[name] = t
        return t[: 2 * count]

    def _pick_victim(self, exclude: _Q) -> _Q:
        locals_ = [o for o in self.q if o.li is not None and o is not exclude]
        if not locals_:
            raise NotImplementedError("swap of a global qubit needs at least one local qubit")
        real = [o for o in locals_ if not self.engine.is_virtual(o.li)]
        pool = real if real else locals_
        if self.victim_priority is not None:
            # the caller knows when each qubit is needed next (e.g. the measurement order of the pattern):
            # evict the one needed farthest in the future (Belady), ties broken towards the newest qubit
            pos = {id(o): i for i, o in enumerate(self.q)}
            return max(pool, key=lambda o: (self.victim_priority(pos[id(o)]), o.li))
        return max(pool, key=lambda o: o.li)       # heuristic: the most recently added local qubit is measured last

    def _swap_in(self, glo: _Q) -> None:
        """Swap global qubit ``glo`` with a local qubit through a pairwise half-shard exchange."""
        if self.profile and self.engine.device.type == "cuda":
            import time  # noqa: PLC0415

            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._swap_in_impl(glo)
            e1.record()
            self._xev.append((e0, e1))
            self.exchange_host_ms += (time.perf_counter() - t0) * 1e3
            return
        self._swap_in_impl(glo)

    def collect_profile(self) -> None:
        if self._xev:
            torch.cuda.synchronize()
            self.exchange_ms += sum(a.elapsed_time(b) for a, b in self._xev)
            self._xev = []

    def _swap_in_impl(self, glo: _Q) -> None:
        k = glo.g
        vic = self._pick_victim(glo)
        self._flush_local(vic)
        self.engine.settle(vic.li)
        B = (self.rank >> k) & 1
        partner = self.rank ^ (1 << k)
        half = self.engine.half_size()
        mag = np.abs(self.u)
        if mag.min() < 1e-100 * mag.max():
            # a rank whose bookkeeping scalar is (numerically) zero — e.g. a global qubit projected onto |0> while
            # every qubit was global — is about to receive amplitudes that are not: fold the scalars into the shards
            self.engine.scale(complex(self.u[self.rank]))
            self.u[:] = 1.0
        factor = complex(self.u[partner] / self.u[self.rank])
        if self._bar is not None:
            # peer-memory path: every rank's earlier kernels are done before anyone touches a partner register, and both
            # in-place swap kernels are done before anyone goes on
            if torch.cuda.current_stream().cuda_stream != self.engine.stream_handle():
                raise RuntimeError("sharded exchange: the current torch stream is not the engine's stream")
            self.engine.flush_batch()      # queued projections touch the halves the partner is about to read
            self._stream_barrier()
            self.engine.swap_peer(vic.li, k, B, factor)
            self._stream_barrier()
            self.exchanged_bytes += 16 * half
            self._after_swap(glo, vic, k)
            return
        # large shards: 4+ chunks per exchange so that packing / unpacking overlap the NVLink transfer; chunks stay
        # between 512 MiB and 2 GiB per staging buffer (four buffers)
        chunk = min(half, max(self.CHUNK_MIN, min(half // 4, self.CHUNK_MAX)))
        # software pipeline over chunks with two staging pairs: the NVLink transfer of chunk c overlaps with
        # packing chunk c+1 and unpacking chunk c-1 (pack/unpack run on the compute stream, NCCL on its own)
        inflight: list[tuple[list, int, int, torch.Tensor]] = []
        for i, first in enumerate(range(0, half, chunk)):
            count = min(chunk, half - first)
            send = self._staging(f"send{i & 1}", count)
            recv = self._staging(f"recv{i & 1}", count)
            # my half with victim bit = 1-B goes to the partner, the partner's half with victim bit = B comes here
            self.engine.pack_half(vic.li, 1 - B, first, count, send)
            ops = [dist.P2POp(dist.isend, send, partner, group=self.group),
                   dist.P2POp(dist.irecv, recv, partner, group=self.group)]
            if self.rank > partner:
                ops.reverse()
            inflight.append((dist.batch_isend_irecv(ops), first, count, recv))
            if len(inflight) == 2:
                reqs, f0, c0, r0_ = inflight.pop(0)
                for req in reqs:
                    req.wait()
                self.engine.unpack_half(vic.li, 1 - B, f0, c0, r0_, factor)
            self.exchanged_bytes += 16 * count
        for reqs, f0, c0, r0_ in inflight:
            for req in reqs:
                req.wait()
            self.engine.unpack_half(vic.li, 1 - B, f0, c0, r0_, factor)
        self._after_swap(glo, vic, k)

    def _after_swap(self, glo: _Q, vic: _Q, k: int) -> None:
        self.n_swaps += 1
        # the victim's slot now carries `glo` with slot value = physical rank bit; undo a pending relabel
        glo.li, glo.g = vic.li, None
        vic.li, vic.g = None, k
        self.owner[k] = vic
        if self.flip[k]:
            self.engine.evolve_single(glo.li, _X)
            self.flip[k] = 0
        # CZs queued between local qubits and `glo` are now local-local
        keep = []
        for a, b in self.pend_lg:
            if b is glo:
                self.engine.entangle(a.li, glo.li)
            else:
                keep.append((a, b))
        self.pend_lg = keep

    # -- relabels, read-out ---------------------------------------------------------------------
    def swap(self, qubits: tuple[int, int]) -> None:
        i, j = qubits
        self._check(i), self._check(j)
        self.q[i], self.q[j] = self.q[j], self.q[i]

    def permute(self, permutation: Sequence[int]) -> None:
        n = len(self.q)
        if len(permutation) != n:
            raise ValueError(f"Permutation has length {len(permutation)}, but {n} qubits expected.")
        if set(permutation) != set(range(n)):
            raise ValueError(f"{permutation} is not a permutation.")
        self.q = [self.q[p] for p in permutation]

    def norm2(self) -> float:
        self.check_deferred()
        for o in self.q:
            if o.li is not None:
                self._flush_local(o)
        w = abs(self.u[self.rank]) ** 2
        (tot,) = self._allreduce([w * self.engine.norm2_local()])
        return tot / (1 << self._dead())

    def flatten(self) -> np.ndarray:
        """Full state on every rank (read-out for tests / small widths): all-gather of the shards."""
        n = len(self.q)
        if n > 28:
            raise MemoryError("flatten() of a sharded state is limited to 28 qubits")
        self.check_deferred()
        for o in self.q:
            if o.li is not None:
                self._flush_local(o)
        local = np.asarray(self.engine.flatten_local(), dtype=np.complex128) * self.u[self.rank]
        nloc = self.engine.nqubit
        if self.P > 1:
            t = torch.from_numpy(np.ascontiguousarray(local).view(np.float64)).to(self.engine.device)
            parts = [torch.empty_like(t) for _ in range(self.P)]
            dist.all_gather(parts, t, group=self.group)
            shards = [p.cpu().numpy().view(np.complex128) for p in parts]
        else:
            shards = [local]
        full = np.zeros((2,) * n, dtype=np.complex128)
        local_axes = [o.li for o in self.q if o.li is not None]     # engine axis of each local qubit, logical order
        for r in range(self.P):
            if any(self.owner[k] is None and (r >> k) & 1 for k in range(self.gbits)):
                continue                                              # replica of a lower rank
            idx = tuple(self._eb(o.g, r) if o.g is not None else slice(None) for o in self.q)
            full[idx] = np.transpose(shards[r].reshape((2,) * nloc), local_axes) if nloc else shards[r][0]
        return full.reshape(1 << n)

    def shard(self) -> np.ndarray:
        """This rank's amplitudes (host copy): the sub-vector selected by this rank's bits of the global qubits, local
        qubits in logical order (first = MSB).  ``layout()`` names the location of every logical qubit and
        ``rank_bit_values()`` the values this rank holds for the global ones; no communication."""
        self.check_deferred()
        for o in self.q:
            if o.li is not None:
                self._flush_local(o)
        local = np.asarray(self.engine.flatten_local(), dtype=np.complex128) * self.u[self.rank]
        nloc = self.engine.nqubit
        if nloc == 0:
            return local.reshape(1)
        local_axes = [o.li for o in self.q if o.li is not None]
        return np.ascontiguousarray(np.transpose(local.reshape((2,) * nloc), local_axes)).reshape(-1)

    def rank_bit_values(self) -> dict[int, int]:
        """Logical index -> value (0/1) that this rank holds for every qubit living on a rank bit."""
        return {i: self._eb(o.g) for i, o in enumerate(self.q) if o.g is not None}

    @property
    def psi(self) -> np.ndarray:
        return self.flatten()

    def fidelity(self, other) -> float:
        inner = np.dot(self.flatten().conjugate(), np.asarray(other.flatten()))
        return float(np.abs(inner) ** 2)

    def isclose(self, other, *, rtol: float = 1e-09, atol: float = 0.0) -> bool:
        return math.isclose(self.fidelity(other), 1, rel_tol=rtol, abs_tol=atol)

    def apply_noise(self, qubits, noise) -> None:  # noqa: ARG002
        raise NoiseNotSupportedError

    def tensor(self, other) -> None:
        """statevec.py:508-524: self <- self (x) other (``other`` is read through ``flatten()``)."""
        vec = np.asarray(other.flatten(), dtype=np.complex128)
        if vec.size == 1:
            self.engine.scale_host(complex(vec[0]))
        elif vec.size == 2:
            self.add_nodes(1, vec)
        else:
            self._add_vector(vec)

    def remove_qubit(self, qubit: int) -> None:
        """statevec.py:417-428."""
        raise NotImplementedError("This method is deprecated. See :meth:`self.project_qubit`.")

    def evolve(self, op, qubits) -> None:
        """statevec.py:361-395.  Only ``Circuit.simulate`` uses multi-qubit operators; pattern
        simulation never does, and the sharded state does not implement them."""
        if len(qubits) == 1:
            self.evolve_single(op, qubits[0])
            return
        raise NotImplementedError("multi-qubit operators are not part of the sharded pattern-simulation path; "
                                  "use B200Statevector.evolve on one GPU")

    def expectation_value(self, op, qubits) -> complex:
        """statevec.py:397-415 (single-qubit operators only, see :meth:`evolve`)."""
        if len(qubits) == 1:
            return self.expectation_single(op, qubits[0])
        raise NotImplementedError("multi-qubit operators are not part of the sharded pattern-simulation path; "
                                  "use B200Statevector.expectation_value on one GPU")

    # read-outs through flatten() (<= 28 qubits), shared with the single-GPU state object — statevec.py:590-739
    to_dict = B200Statevector.to_dict
    to_prob_dict = B200Statevector.to_prob_dict
    _to_dict_map = B200Statevector._to_dict_map
    draw = B200Statevector.draw

    def reset(self) -> None:
        self.check_deferred()
        self.engine.reset()
        self._clear()

    def _renorm(self, extra: Sequence[float] | None = None, set_norm: bool = True):
        """Collect the local norms of the projections queued since the last call, all-reduce them (together with the
        already weighted local values ``extra`` of a synchronous projection, if any), raise the zero-norm RuntimeError of
        any of them, and — unless a synchronous projection is about to do it — set the normalisation of the current
        state.  Returns (norm^2 of the state after the window relative to a normalised start, reduced ``extra``)."""
        if not self._window:
            if extra is None:
                return 1.0, []
            return 1.0, self._allreduce(list(extra))
        import time  # noqa: PLC0415

        t0 = time.perf_counter()
        local = self.engine.take_norms()
        self.sync_wait_ms += (time.perf_counter() - t0) * 1e3
        window, self._window = self._window, []
        if len(local) != len(window):
            raise RuntimeError(f"sharded state: {len(local)} norms for {len(window)} projections")
        # a negative entry is MINUS the analytic ratio (norm after / norm before) of a projection whose norm the engine did
        # not accumulate (outcome probability known to be 1/2, see gxsv_take_norms); the same on every rank
        vals = [rec[0] * n if n >= 0 else 0.0 for rec, n in zip(window, local, strict=True)]
        nw = len(vals)
        red = self._allreduce(vals + list(extra or []))
        tot, red_extra = red[:nw], red[nw:]
        prev = 1.0                       # the window starts from a normalised state (g was set by the last _renorm)
        for i, ((_, big, ratio, phi, assumed, qubit), n) in enumerate(zip(window, tot, strict=True)):
            if local[i] < 0:
                n = tot[i] = prev * (-local[i])
            p = n / prev                 # norm^2 of the kept (larger) row for a normalised input
            n0, n1 = (p * ratio, p) if big else (p, p * ratio)
            if not n0 > 1e-10 and not n1 > 1e-10:
                raise RuntimeError(f"Attempted to remove qubit {qubit} from 0-norm statevector (detected at sync).")
            if big and (n0 > 1e-10) != assumed:
                self.engine.scale_host(phi if n0 > 1e-10 else phi.conjugate())
            prev = n
        if set_norm:
            self.engine.set_norm(tot[-1])
        return prev, red_extra

    def check_deferred(self) -> None:
        self._renorm()

    def sync(self) -> None:
        self.engine.sync()
        self.check_deferred()

    def stats(self):
        return self.engine.stats()

    def reset_stats(self) -> None:
        self.engine.reset_stats()
        self.exchanged_bytes = 0
        self.n_swaps = 0
        self._xev = []
        self.exchange_ms = 0.0
        self.exchange_host_ms = 0.0
        self.sync_wait_ms = 0.0

    def set_option(self, key: int, value: int) -> None:
        self.engine.set_option(key, value)

    def layout(self) -> list[str]:
        return [f"g{o.g}" if o.g is not None else f"l{o.li}" for o in self.q]


class ShardedStatevectorBackend(B200StatevectorBackend):
    """``Backend`` over a :class:`ShardedStatevector`; same contract as ``B200StatevectorBackend``."""

    def __init__(self, state: ShardedStatevector, *, node_index=None, branch_selector=None, symbolic: bool = False) -> None:
        if symbolic:
            raise ValueError("Statevector backend does not support `symbolic` simulation.")
        set_ = object.__setattr__
        set_(self, "state", state)
        set_(self, "node_index", NodeIndex() if node_index is None else node_index)
        set_(self, "branch_selector", RandomBranchSelector() if branch_selector is None else branch_selector)
        set_(self, "symbolic", False)

    @classmethod
    def with_capacity(cls, max_qubits: int, state=None, *, engine=None, group=None, device: int | None = None,
                      fusion: int = 2, strict: bool = True, pattern=None, batch: int = 12,
                      **kwargs) -> "ShardedStatevectorBackend":
        """``max_qubits`` is the TOTAL width (``Pattern.max_space()``); each rank holds ``max_qubits - log2 P`` bits.

        ``pattern`` (optional, like ``TensorNetworkBackend(pattern, ...)``, graphix/sim/tensornet.py:617-674): the
        pattern about to be simulated.  Only its measurement ORDER is used, to decide which local qubit is parked on
        a rank-index bit when a global qubit has to be swapped in (the one measured farthest in the future)."""
        del fusion, state
        P = dist.get_world_size(group) if dist.is_initialized() else 1
        g = P.bit_length() - 1
        if engine is None:
            engine = GpuShardEngine(max(max_qubits - g, 1), device=device, strict=strict, batch=batch)
        backend = cls(ShardedStatevector(engine, group=group, strict=strict), **kwargs)
        if pattern is not None:
            backend.use_schedule(pattern)
        return backend

    def use_schedule(self, pattern) -> None:
        """Record when every node of ``pattern`` is measured (outputs: never) for the eviction policy."""
        when = {}
        for t, cmd in enumerate(pattern):
            if cmd.kind.name == "M":
                when[cmd.node] = t
        ni = self.node_index
        self.state.victim_priority = lambda pos: when.get(ni[pos], float("inf"))

    def measure(self, node: int, measurement, rng=None, *, stacklevel: int = 1):
        """base_backend.py:812-856 through the generic state surface."""
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

    def finalize(self, output_nodes: Sequence[int]) -> None:
        super().finalize(output_nodes)
        self.state.check_deferred()


__all__ = ["GpuShardEngine", "ShardedInput", "ShardedStatevector", "ShardedStatevectorBackend", "BasicStates"]
