"""graphix_b200 — B200-native complex128 statevector backend for graphix patterns.

``B200StatevectorBackend`` is a drop-in ``Backend`` for
``pattern.simulate(backend=...)`` / ``PatternSimulator`` (graphix/simulator.py:396-479);
the amplitudes live in HBM and every command runs as a hand-written sm_100a kernel
behind the C ABI of ``include/gxsv.h``.  There is no CPU fallback.
"""
from .backend import B200StatevectorBackend, NodeIndex
from .statevec import B200Statevector
from .density_matrix import B200DensityMatrix, B200DensityMatrixBackend
from . import mbqc, noise
from .simulator import PatternSimulator

__all__ = ["B200StatevectorBackend", "B200Statevector", "B200DensityMatrix", "B200DensityMatrixBackend", "NodeIndex",
           "PatternSimulator", "mbqc", "noise"]
