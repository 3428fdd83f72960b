"""Host-side mirror of the reference's noise vocabulary for the density-matrix path (BASELINE config 5).

=================================  ===============================================================
here                               reference
=================================  ===============================================================
``KrausData`` / ``KrausChannel``   graphix/channels.py:24-178
``depolarising_channel``           graphix/channels.py:181-200
``two_qubit_depolarising_channel`` graphix/channels.py:223-258
``DepolarisingNoise`` / ``TwoQubitDepolarisingNoise``   graphix/noise_models/depolarising.py:25-76
``ApplyNoise``                     graphix/noise_models/noise_model.py:46-68
``DepolarisingNoiseModel``         graphix/noise_models/depolarising.py:79-149
=================================  ===============================================================

graphix's own objects work with ``B200DensityMatrixBackend`` unchanged (it only needs
``noise.to_krauschannel()`` and an iterable of ``.coef`` / ``.operator``).
"""
from __future__ import annotations

import dataclasses
import enum
import math
from collections.abc import Iterable, Iterator
from typing import ClassVar

import numpy as np

from .mbqc import Ops, ensure_rng


@dataclasses.dataclass(frozen=True)
class KrausData:
    coef: complex
    operator: np.ndarray


class KrausChannel:
    """Completely positive trace-preserving map as a list of Kraus operators (coef * operator)."""

    def __init__(self, kraus_data: Iterable[KrausData]) -> None:
        self._data = list(kraus_data)
        if not self._data:
            raise ValueError("Cannot instantiate the channel with empty data.")
        dim = self._data[0].operator.shape[0]
        self.nqubit = dim.bit_length() - 1
        tot = np.zeros((dim, dim), dtype=np.complex128)
        for k in self._data:
            if k.operator.shape != (dim, dim):
                raise ValueError("All operators must have the same shape.")
            tot += abs(k.coef) ** 2 * (k.operator.conj().T @ k.operator)
        if not np.allclose(tot, np.eye(dim)):
            raise ValueError("The specified channel is not normalized.")

    def __iter__(self) -> Iterator[KrausData]:
        return iter(self._data)

    def __len__(self) -> int:
        return len(self._data)

    def __getitem__(self, i: int) -> KrausData:
        return self._data[i]


def depolarising_channel(prob: float) -> KrausChannel:
    """(1-p) rho + p/3 (X rho X + Y rho Y + Z rho Z) — graphix/channels.py:181-200."""
    return KrausChannel([KrausData(math.sqrt(1 - prob), Ops.I)] +
                        [KrausData(math.sqrt(prob / 3.0), p) for p in (Ops.X, Ops.Y, Ops.Z)])


def two_qubit_depolarising_channel(prob: float) -> KrausChannel:
    """(1-p) rho + p/15 sum over the 15 non-identity two-qubit Paulis — graphix/channels.py:223-258."""
    paulis = (Ops.I, Ops.X, Ops.Y, Ops.Z)
    data = [KrausData(math.sqrt(1 - prob), np.kron(Ops.I, Ops.I))]
    for a in range(4):
        for b in range(4):
            if a or b:
                data.append(KrausData(math.sqrt(prob / 15.0), np.kron(paulis[a], paulis[b])))
    return KrausChannel(data)


@dataclasses.dataclass
class DepolarisingNoise:
    prob: float
    nqubits: ClassVar[int] = 1

    def to_krauschannel(self) -> KrausChannel:
        return depolarising_channel(self.prob)


@dataclasses.dataclass
class TwoQubitDepolarisingNoise:
    prob: float
    nqubits: ClassVar[int] = 2

    def to_krauschannel(self) -> KrausChannel:
        return two_qubit_depolarising_channel(self.prob)


class _NoiseKind(enum.Enum):
    ApplyNoise = enum.auto()


@dataclasses.dataclass
class ApplyNoise:
    """graphix/noise_models/noise_model.py:46-68: applied if ``domain`` is None or has odd outcome parity."""

    noise: object
    nodes: list[int]
    domain: set[int] | None = None
    kind: ClassVar[_NoiseKind] = _NoiseKind.ApplyNoise


class NoiseModel:
    """graphix/noise_models/noise_model.py:75-114."""

    def input_nodes(self, nodes, rng=None, *, stacklevel: int = 1) -> list:
        raise NotImplementedError

    def command(self, cmd, rng=None, *, stacklevel: int = 1) -> list:
        raise NotImplementedError

    def confuse_result(self, cmd, result, rng=None, *, stacklevel: int = 1):
        raise NotImplementedError

    def transpile(self, sequence, rng=None, *, stacklevel: int = 1) -> list:
        return [n_cmd for cmd in sequence for n_cmd in self.command(cmd, rng=rng, stacklevel=stacklevel + 1)]


class DepolarisingNoiseModel(NoiseModel):
    """graphix/noise_models/depolarising.py:79-149."""

    def __init__(self, prepare_error_prob: float = 0.0, x_error_prob: float = 0.0, z_error_prob: float = 0.0,
                 entanglement_error_prob: float = 0.0, measure_channel_prob: float = 0.0,
                 measure_error_prob: float = 0.0) -> None:
        self.prepare_error_prob = prepare_error_prob
        self.x_error_prob = x_error_prob
        self.z_error_prob = z_error_prob
        self.entanglement_error_prob = entanglement_error_prob
        self.measure_error_prob = measure_error_prob
        self.measure_channel_prob = measure_channel_prob

    def input_nodes(self, nodes, rng=None, *, stacklevel: int = 1) -> list:
        return [ApplyNoise(noise=DepolarisingNoise(self.prepare_error_prob), nodes=[node]) for node in nodes]

    def command(self, cmd, rng=None, *, stacklevel: int = 1) -> list:
        k = cmd.kind.name
        if k == "N":
            return [cmd, ApplyNoise(noise=DepolarisingNoise(self.prepare_error_prob), nodes=[cmd.node])]
        if k == "E":
            return [cmd, ApplyNoise(noise=TwoQubitDepolarisingNoise(self.entanglement_error_prob), nodes=list(cmd.nodes))]
        if k == "M":
            return [ApplyNoise(noise=DepolarisingNoise(self.measure_channel_prob), nodes=[cmd.node]), cmd]
        if k == "X":
            return [cmd, ApplyNoise(noise=DepolarisingNoise(self.x_error_prob), nodes=[cmd.node], domain=cmd.domain)]
        if k == "Z":
            return [cmd, ApplyNoise(noise=DepolarisingNoise(self.z_error_prob), nodes=[cmd.node], domain=cmd.domain)]
        if k in ("C", "T", "ApplyNoise"):
            return [cmd]
        raise ValueError("Unexpected signal!")

    def confuse_result(self, cmd, result, rng=None, *, stacklevel: int = 1):
        """Always draws one uniform number per measurement (depolarising.py:141-149)."""
        rng = ensure_rng(rng, stacklevel=stacklevel + 1)
        if rng.uniform() < self.measure_error_prob:
            return 1 - result
        return result
