"""Scan conversion, without a GPU: the device's per-edge clip + per-scanline add_runs (csrc/device/edge_clip.cuh, built
for the host: cb200_debug_loop_runs) against the reference algorithm over whole loops (polygon clip + add_runs,
hpp:2109-2240, restated in oracle/oracle_raster.cpp: oracle_debug_loop_runs).

* Loops inside the canvas: the two run lists are the same MULTISET, bit for bit -- also at 4096^2, where float
  rounding at coordinates in the thousands is what the full-size parity tests are sensitive to.
* Loops that leave the canvas: the inside pieces still give identical runs; the parts beside the canvas are
  projected per edge where the reference lays one boundary segment, so only runs in the boundary columns
  (x = 0, 1, w - 1, w, w + 1: the reference's crossing points sit within an ulp of x = 0 / x = w, not on it) may
  differ, and per pixel they add up to the same coverage up to the rounding of the reference's own boundary lerps
  (a few ulp of the off-canvas coordinates)."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from tests.test_random_scenes import random_scene_wide, integer_scene


def _libs():
    orc, prod = H.oracle_library(), H.product_library()
    orc.oracle_debug_loops.restype = C.c_long
    orc.oracle_debug_loops.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
    orc.oracle_debug_loop_runs.restype = C.c_long
    orc.oracle_debug_loop_runs.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_long]
    return orc, prod


def _loops(orc, frame, draw):
    xy = np.zeros((200000, 2), np.float32)
    counts = np.zeros(20000, np.uint32)
    npts = C.c_long()
    nl = orc.oracle_debug_loops(C.addressof(frame.frame), draw, xy.ctypes.data, len(xy), counts.ctypes.data, len(counts), C.byref(npts))
    assert nl <= len(counts) and npts.value <= len(xy)
    out, at = [], 0
    for l in range(nl):
        out.append(np.ascontiguousarray(xy[at:at + counts[l]]))
        at += int(counts[l])
    return out


def _runs(fn, pts, w, h, cap=1 << 21):
    rxy = np.zeros((cap, 2), np.int32)
    rd = np.zeros(cap, np.float32)
    n = fn(pts.ctypes.data, len(pts), 0.0, 0.0, w, h, rxy.ctypes.data, rd.ctypes.data, cap)
    assert 0 <= n <= cap
    # an edge that starts going up exactly at y = h leaves zero-area runs in row h (add_runs' first, empty scanline);
    # the device only ever emits rows [0, h)
    keep = rxy[:n, 1] < h
    assert not rd[:n][~keep].any()
    return rxy[:n][keep], rd[:n][keep]


def _keys(rxy, rd):
    rd = rd + np.float32(0.0)                        # -0.0 -> +0.0: a zero-area run's sign carries no information
    return np.sort((rxy[:, 1].astype(np.int64) << 48) | (rxy[:, 0].astype(np.int64) << 32) | rd.view(np.uint32).astype(np.int64))


@pytest.mark.parametrize("size,draws", [(512, range(0, 305, 3)), (4096, [0, 1, 7, 40, 101, 102, 103, 150, 222, 304])])
def test_tiger_runs_are_the_reference_runs(size, draws):
    orc, prod = _libs()
    frame = H.lower_script(H.tiger_script(size, size), size, size)[0]
    loops = 0
    for di in draws:
        for pts in _loops(orc, frame, di):
            a = _keys(*_runs(orc.oracle_debug_loop_runs, pts, size, size))
            b = _keys(*_runs(prod.cb200_debug_loop_runs, pts, size, size))
            assert np.array_equal(a, b), "draw %d: run lists differ" % di
            loops += 1
    assert loops >= 10


def _pixel_sums(rxy, rd, w):
    key = rxy[:, 1].astype(np.int64) * (w + 4) + rxy[:, 0].astype(np.int64)
    order = np.argsort(key, kind="stable")
    key, val = key[order], rd[order].astype(np.float64)
    uniq, start = np.unique(key, return_index=True)
    return dict(zip(uniq.tolist(), np.add.reduceat(val, start).tolist())) if len(key) else {}


@pytest.mark.parametrize("generator,seeds", [("wide", list(range(0, 60)) + [51, 672, 679, 695]), ("integer", range(0, 60))])
def test_loops_leaving_the_canvas(generator, seeds):
    orc, prod = _libs()
    crossing = 0
    for seed in seeds:
        script, w, h = random_scene_wide(seed) if generator == "wide" else integer_scene(seed)
        for frame in H.lower_script(script, w, h):
            for di in range(frame.n_draws):
                for pts in _loops(orc, frame, di):
                    ra, rb = _runs(orc.oracle_debug_loop_runs, pts, w, h), _runs(prod.cb200_debug_loop_runs, pts, w, h)
                    a, b = _keys(*ra), _keys(*rb)
                    if np.array_equal(a, b):
                        continue
                    tol = 8.0 * float(np.spacing(np.float32(np.abs(pts).max() + w + h)))
                    # what is left at the end of every scanline (it is painted to the right canvas edge once it
                    # reaches 1/8160, hpp:2573): a loop must not leave more behind than the reference does, beyond
                    # the rounding of the reference's own boundary lerps
                    for y in np.union1d(ra[0][:, 1], rb[0][:, 1]):
                        ea = float(ra[1][ra[0][:, 1] == y].astype(np.float64).sum())
                        eb = float(rb[1][rb[0][:, 1] == y].astype(np.float64).sum())
                        assert abs(ea - eb) <= tol, "seed %d draw %d row %d: residue %.3g, reference %.3g" % (seed, di, y, eb, ea)
                    crossing += 1
                    only = np.concatenate([np.setdiff1d(a, b), np.setdiff1d(b, a)])
                    xs = (only >> 32) & 0xffff
                    assert np.isin(xs, [0, 1, w - 1, w, w + 1]).all(), "seed %d draw %d: runs differ away from the boundary columns" % (seed, di)
                    # The reference clips its boundary segments against the top / bottom edge with a lerp between the
                    # two crossing points (hpp:2220-2222), which lands within a few ulp of the OFF-CANVAS coordinates
                    # of y = 0 / y = h instead of on it; the per-edge projections clamp exactly.  So in the boundary
                    # columns the sums agree to that rounding, not to the bit.
                    sa, sb = _pixel_sums(*ra, w), _pixel_sums(*rb, w)
                    for k in set(sa) | set(sb):
                        assert abs(sa.get(k, 0.0) - sb.get(k, 0.0)) <= tol, (seed, di, k % (w + 4), k // (w + 4), tol)
    assert crossing > 5
