"""Host-side mirror of the reference's vocabulary for the pattern-simulation path.

When graphix itself is importable, its own objects are used with
:class:`graphix_b200.B200StatevectorBackend` unchanged (duck typing: the backend
only touches the attributes listed below).  This module is what drives the same
backend where graphix is absent (the GPU box), with the same names, argument
meaning and error behaviour:

=====================  ==========================================================
here                   reference
=====================  ==========================================================
``Plane.polar``        graphix/fundamentals.py:458-473
``Measurement``        graphix/measurements.py:206-330, 356-440 (Bloch / Pauli)
``PlanarState``        graphix/states.py:40-79, ``BasicStates`` :83-92
``N E M X Z C T``      graphix/command.py:30-41, 60-330
``Clifford``           graphix/clifford.py, table graphix/_db.py:15-66
``Ops``                graphix/ops.py:30-32
``*BranchSelector``    graphix/branch_selector.py:28-157
``Pattern``            graphix/pattern.py (container part only: commands,
                       input/output nodes, ``max_space`` :1343-1370)
=====================  ==========================================================

Angles are in units of pi, as in the reference.
"""
from __future__ import annotations

import abc
import cmath
import dataclasses
import enum
import math
import warnings
from collections.abc import Callable, Iterable, Iterator, Mapping, Sequence
from typing import ClassVar

import numpy as np

Outcome = int


# ---------------------------------------------------------------------------
# planes, measurements
# ---------------------------------------------------------------------------
class Axis(enum.Enum):
    X = 0
    Y = 1
    Z = 2


class Plane(enum.Enum):
    """Measurement plane; ``cos``/``sin`` axes as in graphix/fundamentals.py:404-440."""

    XY = 0
    YZ = 1
    XZ = 2

    def polar(self, angle: float) -> tuple[float, float, float]:
        """Bloch vector at ``angle`` (units of pi) — graphix/fundamentals.py:458-473."""
        rad = angle * math.pi
        c, s = math.cos(rad), math.sin(rad)
        if self is Plane.XY:
            return (c, s, 0.0)
        if self is Plane.YZ:
            return (0.0, s, c)
        return (s, 0.0, c)


@dataclasses.dataclass(frozen=True)
class Measurement:
    """Bloch measurement ``plane(angle)``; Pauli measurements are the six special cases.

    ``Measurement.XY(a)``, ``.YZ(a)``, ``.XZ(a)`` and the constants ``Measurement.X/Y/Z``
    mirror graphix/measurements.py; ``-Measurement.X`` is the opposite sign.
    """

    angle: float
    plane: Plane

    X: ClassVar["Measurement"]
    Y: ClassVar["Measurement"]
    Z: ClassVar["Measurement"]

    @staticmethod
    def XY(angle: float) -> "Measurement":  # noqa: N802
        return Measurement(angle, Plane.XY)

    @staticmethod
    def YZ(angle: float) -> "Measurement":  # noqa: N802
        return Measurement(angle, Plane.YZ)

    @staticmethod
    def XZ(angle: float) -> "Measurement":  # noqa: N802
        return Measurement(angle, Plane.XZ)

    def to_bloch(self) -> "Measurement":
        return self

    def __neg__(self) -> "Measurement":
        return Measurement(_mod2(self.angle + 1.0), self.plane)

    def bloch_vector(self) -> tuple[float, float, float]:
        return self.plane.polar(self.angle)

    # Signal adaptation (graphix/simulator.py:205-228 applies Clifford X for the s-signal and
    # Clifford Z for the t-signal; graphix/measurements.py:312-330 gives the angle rule).
    # X flips the y and z components of the Bloch vector, Z flips x and y.
    def clifford_x(self) -> "Measurement":
        a = self.angle
        if self.plane is Plane.XY:      # (cos, sin, 0) -> (cos, -sin, 0)
            return Measurement(_mod2(-a), self.plane)
        if self.plane is Plane.YZ:      # (0, sin, cos) -> (0, -sin, -cos)
            return Measurement(_mod2(a + 1.0), self.plane)
        return Measurement(_mod2(1.0 - a), self.plane)  # XZ: (sin, 0, cos) -> (sin, 0, -cos)

    def clifford_z(self) -> "Measurement":
        a = self.angle
        if self.plane is Plane.XY:      # (cos, sin, 0) -> (-cos, -sin, 0)
            return Measurement(_mod2(a + 1.0), self.plane)
        return Measurement(_mod2(-a), self.plane)  # YZ: (0,-sin,cos); XZ: (-sin,0,cos)


def _mod2(a: float) -> float:
    a = math.fmod(a, 2.0)
    return a + 2.0 if a < 0 else a


Measurement.X = Measurement(0.0, Plane.XY)
Measurement.Y = Measurement(0.5, Plane.XY)
Measurement.Z = Measurement(0.0, Plane.YZ)


# ---------------------------------------------------------------------------
# states
# ---------------------------------------------------------------------------
class State(abc.ABC):
    @abc.abstractmethod
    def to_statevector_numpy(self) -> np.ndarray: ...


@dataclasses.dataclass(frozen=True)
class PlanarState(State):
    """Single-qubit state in a plane — graphix/states.py:40-79."""

    plane: Plane
    angle: float

    def to_statevector_numpy(self) -> np.ndarray:
        rad = self.angle * math.pi
        if self.plane is Plane.XY:
            return np.asarray([1 / math.sqrt(2), cmath.exp(1j * rad) / math.sqrt(2)], dtype=np.complex128)
        if self.plane is Plane.YZ:
            return np.asarray([math.cos(rad / 2), 1j * math.sin(rad / 2)], dtype=np.complex128)
        return np.asarray([math.cos(rad / 2), math.sin(rad / 2)], dtype=np.complex128)


class BasicStates:
    """graphix/states.py:83-92."""

    ZERO: ClassVar[PlanarState] = PlanarState(Plane.XZ, 0.0)
    ONE: ClassVar[PlanarState] = PlanarState(Plane.XZ, 1.0)
    PLUS: ClassVar[PlanarState] = PlanarState(Plane.XY, 0.0)
    MINUS: ClassVar[PlanarState] = PlanarState(Plane.XY, 1.0)
    PLUS_I: ClassVar[PlanarState] = PlanarState(Plane.XY, 0.5)
    MINUS_I: ClassVar[PlanarState] = PlanarState(Plane.XY, -0.5)
    VEC: ClassVar[list[PlanarState]] = []


BasicStates.VEC = [BasicStates.PLUS, BasicStates.MINUS, BasicStates.ZERO, BasicStates.ONE, BasicStates.PLUS_I,
                   BasicStates.MINUS_I]


# ---------------------------------------------------------------------------
# operators: Pauli and the 24 single-qubit Cliffords
# ---------------------------------------------------------------------------
def _locked(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.complex128)
    a.setflags(write=False)
    return a


_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
_S = np.array([[1, 0], [0, 1j]], dtype=np.complex128)

# Clifford k of the reference's numbering (graphix/_db.py:15-66) written as a word in the
# generators h, s (applied right to left as a matrix product, left to right as text) and a
# global phase exp(i*pi*p/4).  tests/test_mbqc_host.py checks the 24 matrices against the
# golden dump of the reference table.
_CLIFFORD_WORDS: tuple[tuple[str, int], ...] = (
    ("", 0), ("hssh", 0), ("hsshss", 2), ("ss", 0), ("s", 0), ("sss", 0), ("h", 0), ("hsh", 7),
    ("hss", 0), ("hsshs", 5), ("shssh", 5), ("sshss", 0), ("ssh", 4), ("hshss", 1), ("shsss", 2),
    ("shs", 4), ("hsss", 3), ("sshs", 3), ("hshssh", 7), ("hs", 5), ("sh", 5), ("hshs", 2),
    ("hshsss", 0), ("shss", 3),
)


def _word_matrix(word: str, phase8: int) -> np.ndarray:
    m = np.eye(2, dtype=np.complex128)
    for ch in word:
        m = m @ (_H if ch == "h" else _S)
    return cmath.exp(1j * math.pi * phase8 / 4) * m


class Ops:
    """graphix/ops.py:30-32."""

    I: ClassVar[np.ndarray] = _locked(np.eye(2))  # noqa: E741
    X: ClassVar[np.ndarray] = _locked([[0, 1], [1, 0]])
    Y: ClassVar[np.ndarray] = _locked([[0, -1j], [1j, 0]])
    Z: ClassVar[np.ndarray] = _locked([[1, 0], [0, -1]])
    H: ClassVar[np.ndarray] = _locked(_H)
    S: ClassVar[np.ndarray] = _locked(_S)
    CZ: ClassVar[np.ndarray] = _locked(np.diag([1, 1, 1, -1]))


class Clifford:
    """One of the 24 single-qubit Clifford gates, indexed as in graphix/_db.py:41-66."""

    _TABLE: ClassVar[list[np.ndarray] | None] = None
    I: ClassVar["Clifford"]  # noqa: E741
    X: ClassVar["Clifford"]
    Y: ClassVar["Clifford"]
    Z: ClassVar["Clifford"]
    S: ClassVar["Clifford"]
    SDG: ClassVar["Clifford"]
    H: ClassVar["Clifford"]

    def __init__(self, index: int) -> None:
        if not 0 <= int(index) < 24:
            raise ValueError(f"Clifford index {index} out of range [0, 23]")
        self.index = int(index)

    @classmethod
    def table(cls) -> list[np.ndarray]:
        if cls._TABLE is None:
            cls._TABLE = [_locked(_clean(_word_matrix(w, p))) for (w, p) in _CLIFFORD_WORDS]
        return cls._TABLE

    @property
    def matrix(self) -> np.ndarray:
        return self.table()[self.index]

    def __iter__(self) -> Iterator["Clifford"]:  # pragma: no cover - convenience
        return (Clifford(i) for i in range(24))

    def __eq__(self, other: object) -> bool:
        return isinstance(other, Clifford) and other.index == self.index

    def __hash__(self) -> int:
        return hash(("Clifford", self.index))

    def __repr__(self) -> str:
        return f"Clifford({self.index})"


def _clean(m: np.ndarray) -> np.ndarray:
    """Snap entries that are exact up to rounding (0, +-1, +-1/2, +-1/sqrt2 ...) to their values."""
    out = m.copy()
    for v in (0.0, 1.0, -1.0, 0.5, -0.5, 1 / math.sqrt(2), -1 / math.sqrt(2)):
        out.real[np.abs(out.real - v) < 1e-12] = v
        out.imag[np.abs(out.imag - v) < 1e-12] = v
    return out


Clifford.I, Clifford.X, Clifford.Y, Clifford.Z = Clifford(0), Clifford(1), Clifford(2), Clifford(3)
Clifford.S, Clifford.SDG, Clifford.H = Clifford(4), Clifford(5), Clifford(6)


# ---------------------------------------------------------------------------
# commands
# ---------------------------------------------------------------------------
class CommandKind(enum.Enum):
    N = enum.auto()
    M = enum.auto()
    E = enum.auto()
    C = enum.auto()
    X = enum.auto()
    Z = enum.auto()
    S = enum.auto()
    T = enum.auto()


@dataclasses.dataclass
class N:
    node: int
    state: State = dataclasses.field(default_factory=lambda: BasicStates.PLUS)
    kind: ClassVar[CommandKind] = CommandKind.N


@dataclasses.dataclass
class E:
    nodes: tuple[int, int]
    kind: ClassVar[CommandKind] = CommandKind.E


@dataclasses.dataclass
class M:
    node: int
    measurement: Measurement = Measurement.X
    s_domain: set[int] = dataclasses.field(default_factory=set)
    t_domain: set[int] = dataclasses.field(default_factory=set)
    kind: ClassVar[CommandKind] = CommandKind.M


@dataclasses.dataclass
class X:
    node: int
    domain: set[int] = dataclasses.field(default_factory=set)
    kind: ClassVar[CommandKind] = CommandKind.X


@dataclasses.dataclass
class Z:
    node: int
    domain: set[int] = dataclasses.field(default_factory=set)
    kind: ClassVar[CommandKind] = CommandKind.Z


@dataclasses.dataclass
class C:
    node: int
    clifford: Clifford
    kind: ClassVar[CommandKind] = CommandKind.C


@dataclasses.dataclass
class T:
    kind: ClassVar[CommandKind] = CommandKind.T


Command = N | E | M | X | Z | C | T


# ---------------------------------------------------------------------------
# pattern container
# ---------------------------------------------------------------------------
class Pattern:
    """Command list + input/output nodes (the part of graphix.pattern.Pattern the simulator reads)."""

    def __init__(self, input_nodes: Sequence[int] = (), cmds: Iterable[Command] = (),
                 output_nodes: Sequence[int] | None = None) -> None:
        self.input_nodes = list(input_nodes)
        self._cmds: list[Command] = list(cmds)
        self._output_nodes = None if output_nodes is None else list(output_nodes)

    def add(self, cmd: Command) -> None:
        self._cmds.append(cmd)

    def extend(self, cmds: Iterable[Command]) -> None:
        self._cmds.extend(cmds)

    def __iter__(self) -> Iterator[Command]:
        return iter(self._cmds)

    def __len__(self) -> int:
        return len(self._cmds)

    def __getitem__(self, i: int) -> Command:
        return self._cmds[i]

    @property
    def output_nodes(self) -> list[int]:
        """Explicit order if given, else prepared-and-never-measured nodes in preparation order."""
        if self._output_nodes is not None:
            return list(self._output_nodes)
        alive = list(self.input_nodes)
        for c in self._cmds:
            if c.kind is CommandKind.N:
                alive.append(c.node)
            elif c.kind is CommandKind.M:
                alive.remove(c.node)
        return alive

    def max_space(self) -> int:
        """Maximum number of simultaneously live qubits — graphix/pattern.py `max_space`."""
        n = len(self.input_nodes)
        best = n
        for c in self._cmds:
            if c.kind is CommandKind.N:
                n += 1
                best = max(best, n)
            elif c.kind is CommandKind.M:
                n -= 1
        return best

    def check_runnability(self) -> None:
        """Every command acts on live nodes; domains refer to already measured nodes."""
        alive = set(self.input_nodes)
        measured: set[int] = set()
        for c in self._cmds:
            k = c.kind
            if k is CommandKind.N:
                if c.node in alive or c.node in measured:
                    raise ValueError(f"node {c.node} prepared twice")
                alive.add(c.node)
            elif k is CommandKind.E:
                for v in c.nodes:
                    if v not in alive:
                        raise ValueError(f"E on inactive node {v}")
            elif k is CommandKind.M:
                if c.node not in alive:
                    raise ValueError(f"M on inactive node {c.node}")
                if not (c.s_domain | c.t_domain) <= measured:
                    raise ValueError(f"M({c.node}) depends on unmeasured nodes")
                alive.remove(c.node)
                measured.add(c.node)
            elif k in (CommandKind.X, CommandKind.Z):
                if c.node not in alive:
                    raise ValueError(f"{k.name} on inactive node {c.node}")
                if not c.domain <= measured:
                    raise ValueError(f"{k.name}({c.node}) depends on unmeasured nodes")
            elif k is CommandKind.C:
                if c.node not in alive:
                    raise ValueError(f"C on inactive node {c.node}")

    def simulate(self, backend, input_state=BasicStates.PLUS, rng=None, **kwargs):
        """``Pattern.simulate`` — graphix/pattern.py:1439-1478 (backend instances only)."""
        from .simulator import PatternSimulator  # noqa: PLC0415

        sim = PatternSimulator(self, backend=backend, **kwargs)
        sim.run(input_state, rng=rng)
        return sim.backend.state


# ---------------------------------------------------------------------------
# branch selectors
# ---------------------------------------------------------------------------
_default_rng: np.random.Generator | None = None


def ensure_rng(rng: np.random.Generator | None = None, *, stacklevel: int = 1) -> np.random.Generator:
    """graphix/rng.py:17-44."""
    global _default_rng
    if rng is not None:
        return rng
    warnings.warn("Default random-number generator is used. Results may not be reproducible.",
                  stacklevel=stacklevel + 1)
    if _default_rng is None:
        _default_rng = np.random.default_rng()
    return _default_rng


class BranchSelector(abc.ABC):
    """graphix/branch_selector.py:28-58."""

    @abc.abstractmethod
    def measure(self, qubit: int, f_expectation0: Callable[[], float], rng=None, *, stacklevel: int = 1) -> Outcome:
        ...


@dataclasses.dataclass
class RandomBranchSelector(BranchSelector):
    """graphix/branch_selector.py:61-91."""

    pr_calc: bool = True

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        rng = ensure_rng(rng, stacklevel=stacklevel + 1)
        if self.pr_calc:
            prob_0 = f_expectation0()
            return 1 if rng.random() > prob_0 else 0
        return int(rng.choice([0, 1]))


@dataclasses.dataclass
class FixedBranchSelector(BranchSelector):
    """graphix/branch_selector.py:97-135."""

    results: Mapping[int, Outcome]
    default: BranchSelector | None = None

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        result = self.results.get(qubit)
        if result is None:
            if self.default is None:
                raise ValueError(f"Unexpected measurement of qubit {qubit}.")
            return self.default.measure(qubit, f_expectation0)
        return result


@dataclasses.dataclass
class ConstBranchSelector(BranchSelector):
    """graphix/branch_selector.py:138-157."""

    result: Outcome

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        return self.result


class RecordingBranchSelector(BranchSelector):
    """Always evaluates ``f_expectation0`` and records it, then defers to ``inner``.

    Parity-test helper modelled on the reference's CheckedBranchSelector
    (tests/test_branch_selector.py:29-44).
    """

    def __init__(self, inner: BranchSelector) -> None:
        self.inner = inner
        self.p0: dict[int, float] = {}

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        p = f_expectation0()
        self.p0[qubit] = p
        return self.inner.measure(qubit, lambda: p, rng, stacklevel=stacklevel + 1)
