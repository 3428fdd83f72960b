"""Host-side mirror (graphix_b200.mbqc / pattern_io / simulator) against data dumped from the reference."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest
from helpers import GOLDEN, golden_pattern_names, load_golden

from graphix_b200 import mbqc
from graphix_b200.backend import NodeIndex
from graphix_b200.pattern_io import pattern_from_dict, pattern_to_dict


def test_clifford_table_matches_reference():
    g = np.load(os.path.join(GOLDEN, "cliffords.npz"))
    for k in range(24):
        np.testing.assert_allclose(mbqc.Clifford(k).matrix, g["clifford"][k], atol=1e-15)
    for name in ("X", "Y", "Z", "H", "S"):
        np.testing.assert_allclose(getattr(mbqc.Ops, name), g[name], atol=1e-15)
    assert not mbqc.Clifford(3).matrix.flags.writeable


def test_measurement_adaptation_and_states_match_reference():
    d = json.load(open(os.path.join(GOLDEN, "measure_adapt.json")))
    n = 0
    for row in d["rows"]:
        plane = mbqc.Plane[row["plane"]]
        if "state" in row:
            st = mbqc.PlanarState(plane, row["angle"]).to_statevector_numpy()
            np.testing.assert_allclose(st.view(np.float64), row["state"], atol=1e-15)
            continue
        m = mbqc.Measurement(row["angle"], plane)
        if row["s"]:
            m = m.clifford_x()
        if row["t"]:
            m = m.clifford_z()
        np.testing.assert_allclose(m.to_bloch().plane.polar(m.angle), row["vec"], atol=1e-14)
        n += 1
    assert n == 3 * 11 * 4
    for name, vec in d["basic_states"].items():
        np.testing.assert_allclose(getattr(mbqc.BasicStates, name).to_statevector_numpy().view(np.float64), vec,
                                   atol=1e-15)
    for name, ref in d["pauli"].items():
        m = getattr(mbqc.Measurement, name[-1])
        if name.startswith("-"):
            m = -m
        np.testing.assert_allclose(m.bloch_vector(), ref["vec"], atol=1e-15)


@pytest.mark.parametrize("name", golden_pattern_names(with_traces=False))
def test_pattern_roundtrip_and_max_space(name):
    d = load_golden(name)
    p = pattern_from_dict(d)
    assert p.max_space() == d["max_space"]          # graphix Pattern.max_space()
    assert p.output_nodes == d["output_nodes"]
    p.check_runnability()
    d2 = pattern_to_dict(p, name=d["name"])
    assert d2["commands"] == d["commands"]
    assert mbqc.Pattern(p.input_nodes, list(p)).output_nodes is not None


def test_bench_pattern_command_counts():
    """command counts of the BASELINE configs."""
    for name, total in (("config1_rand8x5", 363), ("config2_rand26x20", 3756), ("config3_qaoa29", 8352),
                        ("config4_qft32", 16128), ("config4_qft35", 19320)):
        d = load_golden(name)
        assert len(d["commands"]) == total


def test_node_index():
    ni = NodeIndex()
    ni.extend([5, 7, 9, 11])
    assert ni.index(9) == 2 and ni[1] == 7 and len(ni) == 4
    ni.remove(7)
    assert list(ni) == [5, 9, 11] and ni.index(11) == 2
    ni.swap(0, 2)
    assert list(ni) == [11, 9, 5] and ni.index(5) == 2
    ni.clear()
    assert len(ni) == 0


def test_branch_selectors():
    """graphix/branch_selector.py semantics (reference tests/test_branch_selector.py:47-168)."""
    calls = []

    def f():
        calls.append(1)
        return 0.25

    assert mbqc.ConstBranchSelector(1).measure(0, f) == 1 and not calls
    assert mbqc.FixedBranchSelector({3: 1}).measure(3, f) == 1 and not calls
    with pytest.raises(ValueError):
        mbqc.FixedBranchSelector({}).measure(3, f)
    assert mbqc.FixedBranchSelector({}, default=mbqc.ConstBranchSelector(0)).measure(3, f) == 0
    rng = np.random.default_rng(0)
    outs = [mbqc.RandomBranchSelector().measure(0, f, rng) for _ in range(4000)]
    assert len(calls) == 4000 and abs(np.mean(outs) - 0.75) < 0.03
    calls.clear()
    outs = [mbqc.RandomBranchSelector(pr_calc=False).measure(0, f, rng) for _ in range(4000)]
    assert not calls and abs(np.mean(outs) - 0.5) < 0.03
    with pytest.warns(UserWarning):
        mbqc.RandomBranchSelector().measure(0, f)


def test_check_runnability_errors():
    p = mbqc.Pattern([0], [mbqc.M(1)])
    with pytest.raises(ValueError):
        p.check_runnability()
    p = mbqc.Pattern([0], [mbqc.N(1), mbqc.M(0, mbqc.Measurement.X, {1})])
    with pytest.raises(ValueError):
        p.check_runnability()


def test_ket_string_and_draw_passthrough():
    """``B200Statevector.draw`` (statevec.py:680-739): decimal fallback, graphix's pretty printer when importable."""
    from graphix_b200.statevec import B200Statevector, ket_string

    assert ket_string({"00": 0.70710678, "01": -0.70710678j, "11": 0.5 - 0.5j}) == \
        "0.7071|00> - 0.7071i|01> + (0.5-0.5i)|11>"
    assert ket_string({}) == "0"

    class Fake:
        nqubit = 2

        def to_dict(self, encoding="MSB", *, rtol=0.0, atol=1e-8):  # noqa: ARG002
            return {"00": np.sqrt(0.5) + 0j, "01": np.sqrt(0.5) + 0j}

    text = B200Statevector.draw(Fake())
    assert "00" in text and "01" in text
