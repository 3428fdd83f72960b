"""Host logic of the multi-axis batched pass (k_fused_tile), checked without a GPU: the layout tables that
``gxsv_plan_tile_layouts`` (the planner ``flush_batch`` uses) produces are replayed with the kernel's own address
arithmetic (OR / XOR of table entries selected by thread-index bits and register-slot bits).

Invariants that make the in-place pass and its shared-memory transposes correct:
  * in every layout (thread, register slot) -> physical offset is a bijection onto the tile's index set;
  * in every layout (thread, register slot) -> shared-tile slot is a bijection onto [0, 2^WT), and the slot of an
    amplitude depends only on its physical offset (so a transpose written in layout p is read back in layout p + 1);
  * every stage runs in a layout that holds its axis in the register slot the planner reports (lax), and the stages
    are split over the layouts in order;
  * with the I/O layouts on, the lanes of the first and last layout run over the LOWEST bits of the tile (one contiguous
    512-byte run per warp access) even when batch axes sit on bits 0 / 1.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from graphix_b200 import _build

MAX_LAYOUTS, MAX_TB = 6, 8


def plan(axis_bits, stage_axis, live_mask, R, WT, io):
    _build.build()
    lib = C.CDLL(_build.LIBPATH)
    fn = lib.gxsv_plan_tile_layouts
    fn.restype = C.c_int
    nl = C.c_int(0)
    s_end = (C.c_ubyte * MAX_LAYOUTS)()
    lax = (C.c_ubyte * 16)()
    gbits = (C.c_uint64 * (MAX_LAYOUTS * 16))()
    gtid = (C.c_uint64 * (MAX_LAYOUTS * MAX_TB))()
    sbits = (C.c_ushort * (MAX_LAYOUTS * 16))()
    stid = (C.c_ushort * (MAX_LAYOUTS * MAX_TB))()
    ab = bytes(axis_bits)
    sa = bytes(stage_axis)
    nph = fn(C.c_int(len(axis_bits)), ab, C.c_int(len(stage_axis)), sa, C.c_uint64(live_mask), C.c_int(R), C.c_int(WT),
             C.c_int(io), C.byref(nl), s_end, lax, gbits, gtid, sbits, stid)
    return dict(nph=nph, nl=nl.value, s_end=list(s_end), lax=list(lax), gbits=np.array(gbits).reshape(MAX_LAYOUTS, 16),
                gtid=np.array(gtid).reshape(MAX_LAYOUTS, MAX_TB), sbits=np.array(sbits).reshape(MAX_LAYOUTS, 16),
                stid=np.array(stid).reshape(MAX_LAYOUTS, MAX_TB))


def replay(pl, p, R, NTB):
    """(physical offset, shared slot) of every (thread, register slot) of layout p, as the kernel computes them."""
    offs = np.zeros((1 << NTB, 1 << R), dtype=np.uint64)
    slots = np.zeros((1 << NTB, 1 << R), dtype=np.int64)
    for tg in range(1 << NTB):
        go, so = 0, 0
        for i in range(NTB):
            if (tg >> i) & 1:
                go |= int(pl["gtid"][p, i])
                so ^= int(pl["stid"][p, i])
        for b in range(1 << R):
            offs[tg, b] = go | int(pl["gbits"][p, b])
            slots[tg, b] = so ^ int(pl["sbits"][p, b])
    return offs, slots


def tile_index_set(axis_bits, live_mask, WT):
    F = WT - len(axis_bits)
    dense = [b for b in range(64) if (live_mask >> b) & 1 and b not in axis_bits][:F]
    bits = sorted(set(axis_bits) | set(dense))
    out = []
    for x in range(1 << WT):
        o = 0
        for k, b in enumerate(bits):
            if (x >> k) & 1:
                o |= 1 << b
        out.append(o)
    return bits, sorted(out)


def random_case(rng, R, TW):
    WT = 5 + R + {1: 0, 4: 2, 8: 3}[TW]
    n = int(rng.integers(WT + 1, 34))
    # a register with a hole or two
    live_bits = list(range(n + 2))
    for _ in range(int(rng.integers(0, 3))):
        live_bits.remove(int(rng.choice(live_bits[8:]))) if len(live_bits) > WT + 2 else None
    live_bits = live_bits[:n] if len(live_bits) > n else live_bits
    live_mask = sum(1 << b for b in live_bits)
    A = int(rng.integers(R, min(7, R + 3) + 1))
    if rng.random() < 0.6:
        # force low axes often (the case the I/O layouts exist for)
        low = [b for b in live_bits if b < 6]
        k = min(int(rng.integers(1, 5)), A)
        chosen = list(rng.choice(low, size=min(k, len(low)), replace=False))
        rest = [b for b in live_bits if b not in chosen]
        chosen += list(rng.choice(rest, size=A - len(chosen), replace=False))
    else:
        chosen = list(rng.choice(live_bits, size=A, replace=False))
    axis_bits = [int(b) for b in chosen]
    rng.shuffle(axis_bits)
    nst = int(rng.integers(1, 13))
    # mostly "every stage on a new axis" like the transpiled patterns, sometimes revisiting
    stage_axis = []
    for k in range(nst):
        stage_axis.append(int(rng.integers(A)) if rng.random() < 0.4 else k % A)
    return axis_bits, stage_axis, live_mask, R, WT


@pytest.mark.parametrize("io", [0, 1, 2])
@pytest.mark.parametrize("R,TW", [(4, 1), (4, 4), (4, 8), (3, 1)])
def test_layout_tables_are_consistent(R, TW, io):
    rng = np.random.default_rng(100 * R + 10 * TW + io)
    done = 0
    for _ in range(60):
        axis_bits, stage_axis, live_mask, R_, WT = random_case(rng, R, TW)
        pl = plan(axis_bits, stage_axis, live_mask, R, WT, io)
        assert pl["nph"] >= 1
        if pl["nph"] > 4:
            continue                      # does not fit: flush_batch never queues such a batch
        done += 1
        NTB = WT - R
        bits, want_offsets = tile_index_set(axis_bits, live_mask, WT)
        nl = pl["nl"]
        assert pl["nph"] <= nl <= pl["nph"] + 2
        slot_of = None
        for p in range(nl):
            offs, slots = replay(pl, p, R, NTB)
            assert sorted(int(x) for x in offs.ravel()) == want_offsets, (p, axis_bits)
            assert sorted(int(x) for x in slots.ravel()) == list(range(1 << WT)), (p, axis_bits)
            m = {int(o): int(s) for o, s in zip(offs.ravel(), slots.ravel())}
            if slot_of is None:
                slot_of = m
            else:
                assert m == slot_of, f"layout {p}: an amplitude changes its shared-memory slot"
            # bank groups: the 8 lanes of a quarter warp hit at most 2 slots per bank group (1 = conflict free)
            worst = max(np.bincount(slots[q * 8:(q + 1) * 8, b] & 7, minlength=8).max()
                        for q in range((1 << NTB) // 8) for b in range(1 << R))
            assert worst <= 2
        # stages: in order, each in a layout that keeps its axis in register slot lax
        s_end = pl["s_end"]
        assert s_end[nl - 1] == len(stage_axis)
        s = 0
        for p in range(nl):
            assert s_end[p] >= s
            for k in range(s, s_end[p]):
                la = pl["lax"][k]
                assert int(pl["gbits"][p, 1 << la]) == 1 << axis_bits[stage_axis[k]]
            s = s_end[p]
        # I/O layouts: lanes over the lowest tile bits
        lowest_dense = min(b for b in bits if b not in axis_bits)
        if io == 2 or (io == 1 and lowest_dense >= 3):
            for p in (0, nl - 1):
                assert [int(x) for x in pl["gtid"][p, :NTB]] == [1 << b for b in bits[:NTB]]
    assert done >= 20


def test_plain_shape_needs_no_extra_layouts():
    """7 stages on 7 high axes (the common shape of the transpiled patterns): two phases, no I/O layouts in auto mode,
    lanes on the four dense bits."""
    axis_bits = [20, 9, 13, 25, 17, 30, 11]
    pl = plan(axis_bits, list(range(7)), (1 << 33) - 1, 4, 11, 1)
    assert pl["nph"] == 2 and pl["nl"] == 2
    assert [int(x) for x in pl["gtid"][0, :4]] == [1, 2, 4, 8]
    pl2 = plan([0, 9, 13, 25, 17, 2, 1], list(range(7)), (1 << 33) - 1, 4, 11, 1)      # bits 0 - 2 are all axes
    assert pl2["nph"] == 2 and pl2["nl"] == 4
    assert [int(x) for x in pl2["gtid"][0, :7]] == [1, 2, 4, 8, 16, 32, 64]
    pl3 = plan([2, 15, 28, 3, 16, 29, 4], list(range(7)), (1 << 33) - 1, 4, 11, 1)     # low axes, but bits 0 / 1 are dense
    assert pl3["nph"] == 2 and pl3["nl"] == 2
