"""Parity at BASELINE sizes, where neither the reference nor the oracle can run: size-independent properties.

* config 2 (27 live qubits): the per-command kernels and the lazy batched path — each validated against the reference
  at small widths — must produce the same 26-qubit output state (device fidelity), with unit norm.
* config 4 family (examples/qft_with_tn.py: H on every |+> input gives |0...0>, whose QFT is |+...+>; 30 live qubits,
  13 224 commands): the output must be |+...+> whatever branch is taken (known answer; the reference's own 10- and
  13-qubit traces in tests/golden agree), so P(+) = 1 on every output qubit; run with pseudo-random outcomes so that the X/Z byproduct
  corrections are exercised at full size too.
"""
from __future__ import annotations

import numpy as np
import pytest
from helpers import load_golden

from graphix_b200 import B200StatevectorBackend, PatternSimulator, mbqc
from graphix_b200.pattern_io import pattern_from_dict

pytestmark = pytest.mark.gpu


class SeededOutcomes(mbqc.BranchSelector):
    """Deterministic pseudo-random outcomes that never ask for a probability (flow patterns: every p = 1/2)."""

    def __init__(self, seed: int) -> None:
        self.rng = np.random.default_rng(seed)

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        return int(self.rng.integers(2))


def test_config2_per_command_vs_lazy_batched_same_state():
    pattern = pattern_from_dict(load_golden("config2_rand26x20"))
    states = []
    for kw in ({"fusion": 1}, {"fusion": 2, "strict": False, "batch": 8}):
        b = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=SeededOutcomes(5), **kw)
        PatternSimulator(pattern, b).run()
        assert b.state.nqubit == 26
        assert abs(b.state.norm2() - 1) < 1e-10
        states.append(b.state)
    assert abs(states[0].fidelity(states[1]) - 1) < 1e-10


def test_qft_pattern_known_output_at_30_live_qubits():
    pattern = pattern_from_dict(load_golden("config4_qft29"))
    assert pattern.max_space() == 30 and len(pattern) == 13224
    b = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=SeededOutcomes(11), strict=False,
                                             batch=8)
    sim = PatternSimulator(pattern, b)
    sim.run()
    assert sum(sim.measure_method.results.values()) > 1000          # corrections really were conditional
    plus = np.array([[0.5, 0.5], [0.5, 0.5]], dtype=np.complex128)
    for q in range(b.state.nqubit):
        assert abs(b.state.expectation_single(plus, q) - 1) < 1e-9, q
    assert abs(b.state.norm2() - 1) < 1e-10


# ---------------------------------------------------------------------------------------------------------
# Reference-derived parity at full size: digests of the REAL reference run of BASELINE config 2
# (tests/golden/make_fullsize_golden.py: graphix StatevectorBackend, numba kernels statevec.py:841-1198, 27 live qubits,
# ~17 min per run on 8 cores) — inner products of the 2^26-amplitude output with 4 seeded Gaussian vectors, 1024 sampled
# amplitudes and the norm.  All three execution modes must reproduce them.
# ---------------------------------------------------------------------------------------------------------
def _fullsize_digests():
    import json
    import os

    from helpers import GOLDEN

    with open(os.path.join(GOLDEN, "fullsize", "config2_digests.json")) as f:
        return json.load(f)


_PROBES: dict = {}


def _probe(seed: int, n: int):
    """The seeded Gaussian probe vectors (1 GiB each at 26 qubits) are generated once for all six cases."""
    if (seed, n) not in _PROBES:
        import sys

        from helpers import GOLDEN

        sys.path.insert(0, GOLDEN)
        from make_fullsize_golden import probe_vector

        _PROBES[(seed, n)] = probe_vector(seed, n)
    return _PROBES[(seed, n)]


@pytest.mark.parametrize("mode", ["per_command", "strict_lazy", "batched"])
@pytest.mark.parametrize("run", ["const0", "seeded5"])
def test_config2_against_reference_digests(run, mode):
    import sys

    from helpers import GOLDEN

    sys.path.insert(0, GOLDEN)
    from make_fullsize_golden import sample_indices

    dig = _fullsize_digests()["runs"][run]
    pattern = pattern_from_dict(load_golden("config2_rand26x20"))
    sel = mbqc.FixedBranchSelector({int(k): v for k, v in dig["outcomes"].items()})
    kw = {"per_command": {"fusion": 1}, "strict_lazy": {"fusion": 2}, "batched": {"fusion": 2, "strict": False, "batch": 8}}[mode]
    b = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=sel, **kw)
    PatternSimulator(pattern, b).run()
    flat = b.state.flatten()
    n = dig["nqubit"]
    assert flat.size == 1 << n == 1 << 26
    assert abs(np.vdot(flat, flat).real - dig["norm2"]) < 1e-10
    ref_s = np.asarray(dig["samples"], dtype=np.float64).view(np.complex128)
    got_s = flat[sample_indices(n)]
    # the reference fixes the global phase by which half it keeps (statevec.py:1187-1190): compare up to one phase,
    # taken from the largest sampled amplitude
    k = int(np.argmax(np.abs(ref_s)))
    phase = (ref_s[k] / got_s[k]) / abs(ref_s[k] / got_s[k])
    assert np.max(np.abs(got_s * phase - ref_s)) < 1e-12
    for seed, (re, im) in zip(_fullsize_digests()["probe_seeds"], dig["inner"]):
        z = np.vdot(_probe(seed, n), flat) * phase
        assert abs(z - complex(re, im)) < 1e-9, (seed, z, re, im)       # |z| ~ 1: sum of 2^26 terms
    del flat, b
