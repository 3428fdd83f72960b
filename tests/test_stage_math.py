"""The arithmetic of one batched stage (kernels.cuh: ``stage_pairs`` — the 6-FMA butterfly, its sign tables and its
variants), compiled for the CPU from the kernel source and replayed through ``gxsv_debug_stage`` (no GPU needed).

Reference formulas (kernels.cuh, above ``struct Stage``), with t0 / t1 the two amplitudes of a pair on the stage's
register axis, sq / sv the per-pair signs from the tables tq / tv:
    kind 0:  r0 = t0 + a (sq t1),      r1 = sv (t0 - a (sq t1))
    kind 1:  r0 = a t0 + (sq t1),      r1 = sv (a t0 - (sq t1))
    kind 2:  r0 = m00 t0 + m01 (sq t1), r1 = sv (m10 t0 + m11 (sq t1))
and, as a cross-check of the whole factorisation against the reference projector (statevec.py:1160-1198 semantics):
a kind-0 stage with a = c1 / c0 equals the general stage with the matrix (1, a; 1, -a) of a |+> partner.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from graphix_b200 import _build

_D = C.POINTER(C.c_double)


def _lib():
    _build.build()
    lib = C.CDLL(_build.LIBPATH)
    lib.gxsv_debug_stage.restype = C.c_int
    lib.gxsv_debug_stage.argtypes = [C.c_int] * 6 + [_D, _D, _D, C.c_uint64, C.c_uint64, _D]
    return lib


def run_stage(lib, R, axis, kind, acc, hasv, uniq, v, a, m, tq, tv):
    buf = np.ascontiguousarray(v, dtype=np.complex128).copy()
    a2 = np.array([a.real, a.imag], dtype=np.float64)
    m8 = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(4)).view(np.float64).copy()
    n2 = np.zeros(1)
    rc = lib.gxsv_debug_stage(R, axis, kind, acc, hasv, uniq, buf.view(np.float64).ctypes.data_as(_D), a2.ctypes.data_as(_D),
                              m8.ctypes.data_as(_D), C.c_uint64(tq), C.c_uint64(tv), n2.ctypes.data_as(_D))
    return rc, buf, float(n2[0])


def expected(R, axis, kind, v, a, m, tq, tv, hasv, uniq):
    out = np.array(v, dtype=np.complex128)
    m = np.asarray(m, dtype=np.complex128)
    for lo in range(1 << R):
        if (lo >> axis) & 1:
            continue
        hi = lo | (1 << axis)
        k = ((lo >> (axis + 1)) << axis) | (lo & ((1 << axis) - 1))
        sq = -1.0 if (tq >> (8 * (0 if uniq else k) + 7)) & 1 else 1.0
        sv = (-1.0 if (tv >> (8 * k + 7)) & 1 else 1.0) if hasv else 1.0
        t0, t1 = v[lo], sq * v[hi]
        if kind == 0:
            r0, r1 = t0 + a * t1, sv * (t0 - a * t1)
        elif kind == 1:
            r0, r1 = a * t0 + t1, sv * (a * t0 - t1)
        else:
            r0, r1 = m[0, 0] * t0 + m[0, 1] * t1, sv * (m[1, 0] * t0 + m[1, 1] * t1)
        out[lo], out[hi] = r0, r1
    return out


def rand_table(rng, npairs, uniform=False):
    if uniform:
        return 0x8080808080808080 if rng.random() < 0.5 else 0
    t = 0
    for k in range(npairs):
        if rng.random() < 0.5:
            t |= 0x80 << (8 * k)
    return t


VARIANTS = [  # kind, acc, hasv, uniq — exactly the instantiations of stage_on_axis / stage_by_variant
    (0, 0, 0, 1), (0, 0, 0, 0), (0, 1, 0, 0), (0, 1, 1, 0), (1, 1, 1, 0), (2, 1, 1, 0),
]


@pytest.mark.parametrize("R", [3, 4])
@pytest.mark.parametrize("kind,acc,hasv,uniq", VARIANTS)
def test_stage_variants_against_formulas(R, kind, acc, hasv, uniq):
    lib = _lib()
    rng = np.random.default_rng(1000 * R + 100 * kind + 10 * acc + 2 * hasv + uniq)
    npairs = 1 << (R - 1)
    for axis in range(R):
        for _ in range(20):
            v = rng.normal(size=1 << R) + 1j * rng.normal(size=1 << R)
            a = complex(np.exp(1j * rng.uniform(0, 2 * np.pi))) if rng.random() < 0.7 else complex(rng.normal(), rng.normal())
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            tq = rand_table(rng, npairs, uniform=bool(uniq))
            tv = rand_table(rng, npairs)
            rc, got, n2 = run_stage(lib, R, axis, kind, acc, hasv, uniq, v, a, m, tq, tv)
            assert rc == 0
            want = expected(R, axis, kind, v, a, m, tq, tv, hasv, uniq)
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-14 * (1 + np.abs(want).max()))
            if acc:
                assert abs(n2 - float(np.vdot(want, want).real)) <= 1e-12 * (1 + n2)
            else:
                assert n2 == 0.0


def test_uninstantiated_variant_is_refused():
    lib = _lib()
    rc, _, _ = run_stage(lib, 4, 0, 1, 0, 0, 0, np.zeros(16), 1.0 + 0j, np.eye(2), 0, 0)
    assert rc == -1


def test_factored_stage_equals_general_stage_and_halves_the_norm():
    """kind 0 with a = c1 / c0 is the general stage with rows (1, a), (1, -a) (a |+> partner, the scalar s0 c0 taken out);
    for |a| = 1 the norm^2 after the stage is exactly twice the norm^2 before it (the analytic probability 1/2 that lets
    the host skip the accumulation)."""
    lib = _lib()
    rng = np.random.default_rng(5)
    for axis in range(4):
        v = rng.normal(size=16) + 1j * rng.normal(size=16)
        a = complex(np.exp(1j * rng.uniform(0, 2 * np.pi)))
        tq = rand_table(rng, 8)
        _, fast, n_fast = run_stage(lib, 4, axis, 0, 1, 0, 0, v, a, np.eye(2), tq, 0)
        _, gen, n_gen = run_stage(lib, 4, axis, 2, 1, 1, 0, v, a, np.array([[1, a], [1, -a]]), tq, 0)
        np.testing.assert_allclose(fast, gen, atol=1e-13)
        assert abs(n_fast - n_gen) < 1e-12 * n_gen
        assert abs(n_fast - 2 * float(np.vdot(v, v).real)) < 1e-12 * n_fast
