"""The rollout kernel's pruned nearest-way-point search (csrc/course_index.cuh) must return exactly
the index of the reference's exhaustive scan (calc_target_index,
test_scripts/stanley_controller_ellipse.py:188-212: first minimum over ALL course points).

The search code is __host__ __device__; libsccav_cbf.so exposes it on the CPU through the test hook
sccav_debug_course_index_host, so this runs without a GPU.  The checker is numpy's argmin of the
squared distance computed with the same operations (dx*dx + dy*dy, no FMA)."""
import ctypes as C

import numpy as np
import pytest

from sccav_cbf_b200 import _native as nv
from sccav_cbf_b200.course import config1_course, spline_course


def run(cx, cy, fx, fy, hint=None, dtype=64):
    L = nv.lib()
    cx = np.ascontiguousarray(cx, np.float64); cy = np.ascontiguousarray(cy, np.float64)
    fx = np.ascontiguousarray(fx, np.float64); fy = np.ascontiguousarray(fy, np.float64)
    nq = fx.size
    idx = np.empty(nq, np.int32); full = np.empty(nq, np.int32); ev = np.empty(nq, np.int64)
    hp = None
    if hint is not None:
        hint = np.ascontiguousarray(hint, np.int32)
        hp = hint.ctypes.data
    rc = L.sccav_debug_course_index_host(cx.ctypes.data, cy.ctypes.data, len(cx), fx.ctypes.data, fy.ctypes.data, hp, nq,
                                         dtype, idx.ctypes.data, full.ctypes.data, ev.ctypes.data)
    assert rc == 0
    return idx, full, ev


def numpy_argmin(cx, cy, fx, fy, dt=np.float64):
    cx = cx.astype(dt); cy = cy.astype(dt)
    out = np.empty(fx.size, np.int32)
    for k in range(fx.size):
        dx = dt(fx[k]) - cx; dy = dt(fy[k]) - cy
        out[k] = int(np.argmin(dx * dx + dy * dy))
    return out


def test_pruned_search_equals_exhaustive_scan_on_config1_course():
    cx, cy, _ = config1_course()
    rng = np.random.default_rng(5)
    n = 4000
    # queries near the course (what a rollout sees), with good, stale and absurd hints
    base = rng.integers(0, len(cx), n)
    fx = cx[base] + rng.normal(0, 2.0, n); fy = cy[base] + rng.normal(0, 2.0, n)
    for hint in (base, np.clip(base - 10, 0, None), rng.integers(0, len(cx), n), np.zeros(n, np.int32),
                 np.full(n, 10 ** 6, np.int32), np.full(n, -5, np.int32)):
        idx, full, ev = run(cx, cy, fx, fy, hint)
        assert np.array_equal(idx, full)
        assert np.array_equal(idx, numpy_argmin(cx, cy, fx, fy))
    # with the previous index as hint the search touches a small fraction of the 2034 points
    idx, full, ev = run(cx, cy, fx, fy, np.clip(idx - 10, 0, None))
    assert ev.mean() < 150, ev.mean()
    # far-away and degenerate queries
    fx = rng.uniform(-500, 500, n); fy = rng.uniform(-500, 500, n)
    idx, full, _ = run(cx, cy, fx, fy, rng.integers(0, len(cx), n))
    assert np.array_equal(idx, full) and np.array_equal(idx, numpy_argmin(cx, cy, fx, fy))


def test_ties_pick_the_first_minimum():
    # a course that revisits the same points (figure-of-eight style duplicates) and a circle whose
    # centre is equidistant from every point: the FIRST minimum must win whatever the hint
    t = np.linspace(0, 2 * np.pi, 400, endpoint=False)
    cx = np.concatenate([np.cos(t), np.cos(t), np.cos(t)]) * 8.0
    cy = np.concatenate([np.sin(t), np.sin(t), np.sin(t)]) * 8.0
    rng = np.random.default_rng(1)
    n = 600
    k = rng.integers(0, len(cx), n)
    fx = cx[k] * rng.uniform(0.5, 1.5, n); fy = cy[k] * rng.uniform(0.5, 1.5, n)
    fx[:50] = cx[k[:50]]; fy[:50] = cy[k[:50]]            # exactly on a (triplicated) point
    fx[50:60] = 0.0; fy[50:60] = 0.0                      # the centre
    for hint in (k, rng.integers(0, len(cx), n), np.full(n, len(cx) - 1, np.int32)):
        idx, full, _ = run(cx, cy, fx, fy, hint)
        assert np.array_equal(idx, full)
        assert np.array_equal(idx, numpy_argmin(cx, cy, fx, fy))
    assert (idx[:50] < 400).all()


@pytest.mark.parametrize("P", [1, 2, 15, 16, 17, 127, 128, 129, 1000, 5000])
def test_course_lengths_and_tails(P):
    rng = np.random.default_rng(P)
    s = np.linspace(0, 40, P)
    cx = s * 3.0; cy = 5.0 * np.sin(s / 3.0)
    n = 500
    fx = rng.uniform(-10, 130, n); fy = rng.uniform(-12, 12, n)
    idx, full, _ = run(cx, cy, fx, fy, rng.integers(0, P, n))
    assert np.array_equal(idx, full) and np.array_equal(idx, numpy_argmin(cx, cy, fx, fy))


def test_nan_and_inf_queries_behave_like_argmin():
    cx, cy, _ = config1_course()
    fx = np.array([np.nan, 1.0, np.inf, -np.inf, 1e308]); fy = np.array([0.0, np.nan, 0.0, 1.0, 1e308])
    idx, full, _ = run(cx, cy, fx, fy, np.array([5, 100, 2000, 7, 9], np.int32))
    assert np.array_equal(idx, full)
    assert (idx == 0).all()


def test_fp32_variant_matches_its_own_exhaustive_scan():
    cx, cy, _ = spline_course([0.0, 30.0, 60.0, 40.0], [0.0, 10.0, -5.0, -30.0], ds=0.05)
    rng = np.random.default_rng(11)
    n = 3000
    base = rng.integers(0, len(cx), n)
    fx = cx[base] + rng.normal(0, 1.5, n); fy = cy[base] + rng.normal(0, 1.5, n)
    idx, full, _ = run(cx, cy, fx, fy, np.clip(base - 8, 0, None), dtype=32)
    assert np.array_equal(idx, full)


@pytest.mark.parametrize("shift", [0.0, 3.0e3, 5.0e5, 4.0e6])
def test_single_precision_bounds_stay_conservative(shift):
    """The capsule tests run in fp32 relative to an origin on the course with every rounding error charged to the
    slack (course_index.cuh).  Courses far from the coordinate origin (map / UTM coordinates), queries from
    centimetres to 1e7 m away, near-ties on arcs seen from their centre of curvature, and exact duplicates."""
    rng = np.random.default_rng(int(shift) % 97)
    cx, cy, _ = config1_course()
    cx = cx + shift; cy = cy - 0.5 * shift
    n = 3000
    base = rng.integers(0, len(cx), n)
    for spread in (0.02, 2.0, 50.0, 3.0e3, 1.0e7):
        fx = cx[base] + rng.normal(0, spread, n); fy = cy[base] + rng.normal(0, spread, n)
        for hint in (base, rng.integers(0, len(cx), n)):
            idx, full, _ = run(cx, cy, fx, fy, hint)
            assert np.array_equal(idx, full), (shift, spread)
    # an arc of constant radius seen from (almost) its centre: every point is a near-tie
    t = np.linspace(0.0, 1.5 * np.pi, 1500)
    ax = shift + 25.0 * np.cos(t); ay = 25.0 * np.sin(t)
    fx = shift + rng.normal(0, 1e-3, 400); fy = rng.normal(0, 1e-3, 400)
    idx, full, _ = run(ax, ay, fx, fy, rng.integers(0, len(ax), 400))
    assert np.array_equal(idx, full)
    idx, full, _ = run(ax, ay, fx, fy, rng.integers(0, len(ax), 400), dtype=32)
    assert np.array_equal(idx, full)


def test_closed_loop_queries_cost_a_few_dozen_evaluations():
    """Hints as the rollout kernel makes them (previous index + previous advance) along a sweep of the course."""
    cx, cy, cyaw = config1_course()
    k = np.arange(0, len(cx) - 1, 8)
    off = 1.5 * np.sin(k / 40.0)
    fx = cx[k] - off * np.sin(cyaw[k]); fy = cy[k] + off * np.cos(cyaw[k])
    idx0, full, _ = run(cx, cy, fx, fy, k)
    assert np.array_equal(idx0, full)
    hint = np.concatenate([[0, idx0[0]], idx0[1:-1] + (idx0[1:-1] - idx0[:-2])])
    idx, full, ev = run(cx, cy, fx, fy, hint)
    assert np.array_equal(idx, full)
    assert ev.mean() < 60, ev.mean()


def _cover(P):
    L = nv.lib()
    nleaf = C.c_int32(); nlev = C.c_int32(); kc = C.c_int32()
    assert L.sccav_debug_cover_host(P, C.addressof(nleaf), C.addressof(nlev), C.addressof(kc), None, None, None) == 0
    cnt = np.empty(nleaf.value, np.int32)
    lev = np.empty((nleaf.value, kc.value), np.int32); idx = np.empty((nleaf.value, kc.value), np.int32)
    assert L.sccav_debug_cover_host(P, C.addressof(nleaf), C.addressof(nlev), C.addressof(kc), cnt.ctypes.data,
                                    lev.ctypes.data, idx.ctypes.data) == 0
    return nleaf.value, nlev.value, kc.value, cnt, lev, idx


@pytest.mark.parametrize("P", [1, 8, 9, 16, 17, 24, 25, 63, 64, 65, 255, 1000, 2034, 2360, 4999, 8191])
def test_cover_is_a_partition_that_grows_with_distance(P):
    """The search proves its answer by excluding a COVER of everything outside the two-leaf window
    (course_index.cuh, cover_node / cover_row).  For every window leaf w of a course of P points:
    the cover nodes and the window tile the leaves [0, nleaf) exactly -- no leaf missing, none twice --,
    the row fits the table (count <= kc, kc = 3 (nlev - 1) + 1 rounded up to 8), the top level is never used,
    and a node of 2^h leaves lies at least 2^h - 1 leaves from the window (never large where it is close)."""
    nleaf, nlev, kc, cnt, lev, idx = _cover(P)
    assert nleaf == (P + 7) // 8 and kc % 8 == 0 and kc >= 3 * (nlev - 1) + 1
    for w in range(max(nleaf - 1, 1)):
        win = (w, min(w + 2, nleaf))
        seen = np.zeros(nleaf, np.int32)
        seen[win[0]:win[1]] += 1
        assert 0 <= cnt[w] <= kc
        assert (lev[w, cnt[w]:] == -1).all()
        last_level = -1
        for h, j in zip(lev[w, :cnt[w]], idx[w, :cnt[w]]):
            assert 0 <= h < max(nlev - 1, 1)
            assert h >= last_level                      # nearest (smallest) first
            last_level = h
            lo, hi = j << h, min((j + 1) << h, nleaf)
            assert lo < hi
            seen[lo:hi] += 1
            gap = win[0] - hi if hi <= win[0] else lo - win[1]
            assert gap >= (1 << h) - 1, (w, h, j, gap)
        assert (seen == 1).all(), (P, w, np.flatnonzero(seen != 1)[:8])
