"""numpy stand-in for the local shard engine (``graphix_b200.sharded.GpuShardEngine``).

TEST DOUBLE ONLY: lets the host-side sharding logic of ``graphix_b200/sharded.py`` (layout decisions,
exchange schedule, collectives) run under ``gloo`` on CPU with world_size > 1.  It keeps the same
contract as the CUDA engine: projections return LOCAL half-norms and leave the state unnormalised
until ``set_norm(total)``; freshly added qubits stay virtual until something other than a diagonal
gate touches the engine.
"""
from __future__ import annotations

import numpy as np
import torch


class CpuShardEngine:
    def __init__(self, local_bits: int) -> None:
        self.capacity = local_bits
        self.device = torch.device("cpu")
        self.reset()

    def reset(self) -> None:
        self.psi = np.array([1], dtype=np.complex128)   # canonical order over materialised local qubits
        self.n_real = 0
        self.virt: list[list[complex]] = []             # trailing virtual qubits

    # -- queries
    @property
    def nqubit(self) -> int:
        return self.n_real + len(self.virt)

    @property
    def nreal(self) -> int:
        return self.n_real

    def is_virtual(self, q: int) -> bool:
        return q >= self.n_real

    def half_size(self) -> int:
        return 1 << (self.n_real - 1)

    def _materialise(self) -> None:
        for s0, s1 in self.virt:
            self.psi = np.kron(self.psi, np.array([s0, s1], dtype=np.complex128))
            self.n_real += 1
        self.virt = []

    def _t(self) -> np.ndarray:
        return self.psi.reshape((2,) * self.n_real)

    # -- commands
    def add_qubit(self, s0: complex, s1: complex) -> None:
        assert self.nqubit < self.capacity
        self.virt.append([s0, s1])

    def entangle(self, a: int, b: int) -> None:
        self._materialise()
        t = self._t()
        idx = [slice(None)] * self.n_real
        idx[a], idx[b] = 1, 1
        t[tuple(idx)] *= -1

    def evolve_single(self, q: int, m) -> None:
        m = np.asarray(m, dtype=np.complex128)
        if q >= self.n_real and m[0, 1] == 0 and m[1, 0] == 0:
            s = self.virt[q - self.n_real]
            s[0], s[1] = m[0, 0] * s[0], m[1, 1] * s[1]
            return
        self._materialise()
        t = np.moveaxis(self._t(), q, 0)
        new = np.tensordot(m, t, axes=(1, 0))
        self.psi = np.ascontiguousarray(np.moveaxis(new, 0, q)).reshape(-1)

    def expectation_local(self, q: int, m) -> complex:
        self._materialise()
        m = np.asarray(m, dtype=np.complex128)
        t = np.moveaxis(self._t(), q, 0)
        new = np.tensordot(m, t, axes=(1, 0))
        return complex(np.vdot(t, new))

    def project_local(self, q: int, m) -> tuple[float, float]:
        self._materialise()
        m = np.asarray(m, dtype=np.complex128)
        t = np.moveaxis(self._t(), q, 0)
        rows = np.tensordot(m, t, axes=(1, 0))
        n0 = float(np.vdot(rows[0], rows[0]).real)
        n1 = float(np.vdot(rows[1], rows[1]).real)
        r0 = abs(m[0, 0]) ** 2 + abs(m[0, 1]) ** 2
        r1 = abs(m[1, 0]) ** 2 + abs(m[1, 1]) ** 2
        big = 0 if r0 >= r1 else 1
        self.psi = np.ascontiguousarray(rows[big]).reshape(-1)
        self.n_real -= 1
        return n0, n1

    def set_norm(self, total: float) -> None:
        self.psi = self.psi / np.sqrt(total)

    def scale_host(self, c: complex) -> None:
        self.psi = self.psi * c

    def scale(self, c: complex) -> None:
        self._materialise()
        self.psi = self.psi * c

    def settle(self, q: int) -> None:
        self._materialise()

    def _half_view(self, q: int, bit: int) -> np.ndarray:
        return self.psi.reshape(1 << q, 2, 1 << (self.n_real - 1 - q))[:, bit, :]

    def pack_half(self, q: int, bit: int, first: int, count: int, out: torch.Tensor) -> None:
        self._materialise()
        flat = np.ascontiguousarray(self._half_view(q, bit)).reshape(-1)
        out.numpy()[:count] = flat[first:first + count]

    def unpack_half(self, q: int, bit: int, first: int, count: int, src: torch.Tensor, factor: complex) -> None:
        self.psi = np.ascontiguousarray(self.psi)
        view = self._half_view(q, bit)
        flat = np.ascontiguousarray(view).reshape(-1)
        flat[first:first + count] = src.numpy()[:count] * factor
        view[...] = flat.reshape(view.shape)

    def norm2_local(self) -> float:
        self._materialise()
        return float(np.vdot(self.psi, self.psi).real)

    def flatten_local(self) -> np.ndarray:
        self._materialise()
        return self.psi.copy()

    def sync(self) -> None:
        pass

    def stats(self):
        return {}

    def reset_stats(self) -> None:
        pass

    def set_option(self, key: int, value: int) -> None:
        pass
