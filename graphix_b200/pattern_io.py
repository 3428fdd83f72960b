"""(De)serialisation of MBQC patterns as small JSON documents.

graphix is not installed on the GPU box, so the benchmark / parity patterns are
generated once from the reference generators (``tests/golden/make_golden.py``,
run where ``/root/reference`` exists) and stored under ``tests/golden/``.
``pattern_to_dict`` duck-types both graphix ``Pattern`` objects and
``graphix_b200.mbqc.Pattern``; ``pattern_from_dict`` always builds the latter.

Command encoding (angles in units of pi):
``["N", node]`` (|+>) or ``["N", node, plane, angle]``; ``["E", a, b]``;
``["M", node, plane, angle, [s_domain], [t_domain]]``; ``["X"|"Z", node, [domain]]``;
``["C", node, clifford_index]``; ``["T"]``.
"""
from __future__ import annotations

import gzip
import json
import os
from typing import Any

from . import mbqc


def pattern_to_dict(pattern, name: str = "", meta: dict[str, Any] | None = None) -> dict[str, Any]:
    cmds: list[list[Any]] = []
    for c in pattern:
        k = c.kind.name
        if k == "N":
            st = c.state
            plane, angle = st.plane.name, float(st.angle)
            cmds.append(["N", int(c.node)] if (plane == "XY" and angle == 0.0) else ["N", int(c.node), plane, angle])
        elif k == "E":
            cmds.append(["E", int(c.nodes[0]), int(c.nodes[1])])
        elif k == "M":
            b = c.measurement.to_bloch()
            cmds.append(["M", int(c.node), b.plane.name, float(b.angle), sorted(int(x) for x in c.s_domain),
                         sorted(int(x) for x in c.t_domain)])
        elif k in ("X", "Z"):
            cmds.append([k, int(c.node), sorted(int(x) for x in c.domain)])
        elif k == "C":
            cmds.append(["C", int(c.node), int(getattr(c.clifford, "index", getattr(c.clifford, "value", 0)))])
        elif k == "T":
            cmds.append(["T"])
        else:
            raise ValueError(f"cannot serialise command kind {k}")
    d = {
        "name": name,
        "input_nodes": [int(x) for x in pattern.input_nodes],
        "output_nodes": [int(x) for x in pattern.output_nodes],
        "max_space": int(pattern.max_space()),
        "commands": cmds,
    }
    if meta:
        d["meta"] = meta
    return d


def pattern_from_dict(d: dict[str, Any]) -> mbqc.Pattern:
    cmds: list[mbqc.Command] = []
    for c in d["commands"]:
        k = c[0]
        if k == "N":
            if len(c) == 2:
                cmds.append(mbqc.N(c[1]))
            else:
                cmds.append(mbqc.N(c[1], mbqc.PlanarState(mbqc.Plane[c[2]], c[3])))
        elif k == "E":
            cmds.append(mbqc.E((c[1], c[2])))
        elif k == "M":
            cmds.append(mbqc.M(c[1], mbqc.Measurement(c[3], mbqc.Plane[c[2]]), set(c[4]), set(c[5])))
        elif k == "X":
            cmds.append(mbqc.X(c[1], set(c[2])))
        elif k == "Z":
            cmds.append(mbqc.Z(c[1], set(c[2])))
        elif k == "C":
            cmds.append(mbqc.C(c[1], mbqc.Clifford(c[2])))
        elif k == "T":
            cmds.append(mbqc.T())
        else:
            raise ValueError(f"unknown command {c!r}")
    return mbqc.Pattern(d["input_nodes"], cmds, d["output_nodes"])


def save_pattern(path: str, d: dict[str, Any]) -> None:
    data = json.dumps(d, separators=(",", ":")).encode()
    if path.endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def load_pattern_dict(path: str) -> dict[str, Any]:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        return json.loads(f.read().decode())


def load_pattern(path: str) -> mbqc.Pattern:
    return pattern_from_dict(load_pattern_dict(path))


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
