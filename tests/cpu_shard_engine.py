"""numpy stand-in for the local shard engine (``graphix_b200.sharded.GpuShardEngine``).

TEST DOUBLE ONLY: lets the host-side sharding logic of ``graphix_b200/sharded.py`` (layout decisions,
exchange schedule, collectives) run under ``gloo`` on CPU with world_size > 1.  It keeps the contract
of the CUDA engine in distributed mode, including its laziness: a prepared qubit stays *virtual*
(a 2-vector plus queued CZs) until a non-diagonal operation touches it or one of its CZ partners;
projections return LOCAL half-norms and leave the state unnormalised until ``set_norm(total)``.
"""
from __future__ import annotations

import numpy as np
import torch


class CpuShardEngine:
    def __init__(self, local_bits: int, strict: bool = True) -> None:
        self.capacity = local_bits
        self.device = torch.device("cpu")
        self.strict = strict
        self._norms: list[float] = []
        self.reset()

    def reset(self) -> None:
        self.ids: list[int] = []                 # logical order -> stable id
        self.virt: dict[int, list[complex]] = {}  # id -> prepared state while virtual
        self.real: list[int] = []                # ids of materialised qubits = axes of self.t, in order
        self.t = np.ones((), dtype=np.complex128)
        self.pend: list[tuple[int, int]] = []    # queued CZs (ids), at least one endpoint virtual
        self._next = 0

    # -- queries
    @property
    def nqubit(self) -> int:
        return len(self.ids)

    @property
    def nreal(self) -> int:
        return len(self.real)

    def is_virtual(self, q: int) -> bool:
        return self.ids[q] in self.virt

    def half_size(self) -> int:
        return 1 << (len(self.real) - 1)

    # -- laziness
    def _materialise(self, i: int) -> None:
        if i not in self.virt:
            return
        s = np.array(self.virt.pop(i), dtype=np.complex128)
        self.t = np.multiply.outer(self.t, s)
        self.real.append(i)
        keep = []
        for a, b in self.pend:
            if a not in self.virt and b not in self.virt:
                self._cz(a, b)
            else:
                keep.append((a, b))
        self.pend = keep

    def _cz(self, a: int, b: int) -> None:
        idx = [slice(None)] * len(self.real)
        idx[self.real.index(a)], idx[self.real.index(b)] = 1, 1
        self.t[tuple(idx)] *= -1

    def _settle_id(self, i: int) -> None:
        self._materialise(i)
        for a, b in list(self.pend):
            if a == i or b == i:
                self._materialise(b if a == i else a)

    def _all(self) -> None:
        for i in list(self.ids):
            self._materialise(i)

    def _axis(self, q: int) -> int:
        return self.real.index(self.ids[q])

    # -- commands
    def add_qubit(self, s0: complex, s1: complex) -> None:
        assert self.nqubit < self.capacity
        self.ids.append(self._next)
        self.virt[self._next] = [s0, s1]
        self._next += 1

    def tensor(self, vec) -> None:
        vec = np.asarray(vec, dtype=np.complex128).reshape(-1)
        k = int(vec.size).bit_length() - 1
        assert self.nqubit + k <= self.capacity
        self._all()
        self.t = np.multiply.outer(self.t, vec.reshape((2,) * k))
        for _ in range(k):
            self.ids.append(self._next)
            self.real.append(self._next)
            self._next += 1

    def entangle(self, a: int, b: int) -> None:
        ia, ib = self.ids[a], self.ids[b]
        if ia in self.virt or ib in self.virt:
            self.pend.append((ia, ib))
        else:
            self._cz(ia, ib)

    def evolve_single(self, q: int, m) -> None:
        m = np.asarray(m, dtype=np.complex128)
        i = self.ids[q]
        diagonal = m[0, 1] == 0 and m[1, 0] == 0
        if i in self.virt and (diagonal or not any(i in e for e in self.pend)):
            s = self.virt[i]
            s[0], s[1] = m[0, 0] * s[0] + m[0, 1] * s[1], m[1, 0] * s[0] + m[1, 1] * s[1]
            return
        if i in self.virt or not diagonal:
            self._settle_id(i)
        ax = self._axis(q)
        self.t = np.moveaxis(np.tensordot(m, np.moveaxis(self.t, ax, 0), axes=(1, 0)), 0, ax)

    def expectation_local(self, q: int, m) -> complex:
        self._settle_id(self.ids[q])
        m = np.asarray(m, dtype=np.complex128)
        t = np.moveaxis(self.t, self._axis(q), 0)
        return complex(np.vdot(t, np.tensordot(m, t, axes=(1, 0))))

    def project_local(self, q: int, m) -> tuple[float, float]:
        i = self.ids[q]
        self._settle_id(i)
        m = np.asarray(m, dtype=np.complex128)
        ax = self._axis(q)
        rows = np.tensordot(m, np.moveaxis(self.t, ax, 0), axes=(1, 0))
        n0 = float(np.vdot(rows[0], rows[0]).real)
        n1 = float(np.vdot(rows[1], rows[1]).real)
        r0 = abs(m[0, 0]) ** 2 + abs(m[0, 1]) ** 2
        r1 = abs(m[1, 0]) ** 2 + abs(m[1, 1]) ** 2
        self.t = np.array(rows[0 if r0 >= r1 else 1], dtype=np.complex128)  # keeps 0-d for the scalar state
        self.real.remove(i)
        self.ids.remove(i)
        return n0, n1

    def project_local_async(self, q: int, m) -> None:
        m = np.asarray(m, dtype=np.complex128)
        n0, n1 = self.project_local(q, m)
        r0 = abs(m[0, 0]) ** 2 + abs(m[0, 1]) ** 2
        r1 = abs(m[1, 0]) ** 2 + abs(m[1, 1]) ** 2
        self._norms.append(n0 if r0 >= r1 else n1)

    def take_norms(self) -> list[float]:
        out, self._norms = self._norms, []
        return out

    def set_norm(self, total: float) -> None:
        self.t = np.asarray(self.t / np.sqrt(total))

    def scale_host(self, c: complex) -> None:
        self.t = np.asarray(self.t * c)

    def scale(self, c: complex) -> None:
        self.t = np.asarray(self.t * c)

    def settle(self, q: int) -> None:
        self._settle_id(self.ids[q])

    def _half_view(self, q: int, bit: int) -> np.ndarray:
        ax = self._axis(q)
        self.t = np.array(self.t, dtype=np.complex128, order="C")
        n = len(self.real)
        return self.t.reshape(1 << ax, 2, 1 << (n - 1 - ax))[:, bit, :]

    def pack_half(self, q: int, bit: int, first: int, count: int, out: torch.Tensor) -> None:
        i = self.ids[q]
        assert i not in self.virt and not any(i in e for e in self.pend), "qubit must be settled first"
        flat = np.ascontiguousarray(self._half_view(q, bit)).reshape(-1)
        out.numpy().view(np.complex128)[:count] = flat[first:first + count]

    def unpack_half(self, q: int, bit: int, first: int, count: int, src: torch.Tensor, factor: complex) -> None:
        view = self._half_view(q, bit)
        flat = np.ascontiguousarray(view).reshape(-1)
        flat[first:first + count] = src.numpy().view(np.complex128)[:count] * factor
        view[...] = flat.reshape(view.shape)

    def norm2_local(self) -> float:
        self._all()
        return float(np.vdot(self.t, self.t).real)

    def flatten_local(self) -> np.ndarray:
        self._all()
        order = [self.real.index(i) for i in self.ids]
        return np.ascontiguousarray(np.transpose(self.t, order)).reshape(-1)

    def sync(self) -> None:
        pass

    def stats(self):
        return {}

    def reset_stats(self) -> None:
        pass

    def set_option(self, key: int, value: int) -> None:
        pass
