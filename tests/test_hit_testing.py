"""Bulk hit testing: is_point_in_path (hpp:3101-3132) for many points in one launch.

Chain of evidence, every link exact (bools):
  reference's own is_point_in_path, n calls (oracle/_ref, cv_points_in_path of the reference build)
    == oracle_points_in_path over the front end's flattened edges (cv_path_edges)      [CPU, pins the oracle
       and the host flattening against the real reference]
    == cb200_hit_test / cv_points_in_path on the GPU                                   [-m gpu]
Query sets mix random points with lattice points that sit exactly on edges and vertices (the
`side == 0` and horizontal-edge rules) and points on scanlines through vertices (the half-open rule)."""
import ctypes as C
import math

import numpy as np
import pytest

from tests import harness as H


def _star(w, cx, cy, r, n=5, step=2):
    """Self-intersecting star: winding 2 in the core, 1 in the arms."""
    for k in range(n + 1):
        a = 2 * math.pi * ((k * step) % n) / n - math.pi / 2
        w.floats("MOVE_TO" if k == 0 else "LINE_TO", cx + r * math.cos(a), cy + r * math.sin(a))
    w.bare("CLOSE_PATH")


def scene(kind):
    w = H.ScriptWriter()
    if kind == "rects_and_hole":
        w.floats("RECTANGLE", 32, 32, 160, 128)             # integer edges: lattice queries land exactly on them
        w.floats("MOVE_TO", 64, 64)                         # opposite winding: a hole
        w.floats("LINE_TO", 64, 128); w.floats("LINE_TO", 128, 128); w.floats("LINE_TO", 128, 64)
        w.bare("CLOSE_PATH")
        w.floats("RECTANGLE", 100, 100, 120, 120)           # overlaps both
    elif kind == "star":
        _star(w, 128, 128, 110)
        _star(w, 60, 200, 40, n=7, step=3)
    elif kind == "curves":
        w.floats("MOVE_TO", 20, 200)
        w.floats("BEZIER_CURVE_TO", 20, -80, 236, 330, 236, 40)
        w.floats("QUADRATIC_CURVE_TO", 128, 128, 200, 220)
        w.floats("ARC", 128, 128, 70, 0.3, 4.9, 0)
        w.floats("ARC_TO", 10, 10, 250, 30, 40)
        w.floats("MOVE_TO", 10, 10); w.floats("LINE_TO", 240, 16); w.floats("LINE_TO", 128, 250)    # open subpath: closed implicitly
    elif kind == "transformed":
        w.floats("TRANSLATE", 128, 128); w.floats("ROTATE", 0.5); w.floats("SCALE", 1.5, 0.75)
        w.floats("RECTANGLE", -60, -60, 120, 120)
        w.floats("ARC", 0, 0, 50, 0, 6.2831855, 1)
        _star(w, 10, -10, 70)
    elif kind == "long_path":                               # thousands of edges: the edge-chunk grid
        state = [777]

        def u():
            state[0] = (state[0] * 1664525 + 1013904223) & 0xffffffff
            return (state[0] >> 8) / float(1 << 24)
        for _ in range(40):
            w.floats("MOVE_TO", 256 * u(), 256 * u())
            for _ in range(30):
                w.floats("BEZIER_CURVE_TO", *[256 * u() for _ in range(6)])
            w.bare("CLOSE_PATH")
    elif kind == "degenerate":
        w.floats("MOVE_TO", 50, 50)                         # a lone point, then horizontal and vertical slivers
        w.floats("MOVE_TO", 10, 100); w.floats("LINE_TO", 200, 100); w.floats("LINE_TO", 100, 100)
        w.floats("MOVE_TO", 120, 10); w.floats("LINE_TO", 120, 200)
    elif kind == "empty":
        pass
    else:
        raise ValueError(kind)
    return w.take()


SCENES = ["rects_and_hole", "star", "curves", "transformed", "long_path", "degenerate", "empty"]


def queries(n_random, seed=1):
    rng = np.random.default_rng(seed)
    rnd = (rng.random((n_random, 2), dtype=np.float32) * np.float32(300.0) - np.float32(22.0))
    gx, gy = np.meshgrid(np.arange(0, 260, 4, dtype=np.float32), np.arange(0, 260, 4, dtype=np.float32))
    lattice = np.stack([gx.ravel(), gy.ravel()], axis=1)     # hits integer edges / vertices exactly
    half = lattice + np.float32(0.5)
    return np.ascontiguousarray(np.concatenate([rnd, lattice, half]).astype(np.float32))


def _path_on(lib, script, make):
    h = make()
    assert h
    H._run(lib, h, script)
    return h


def _edges(prod, script):
    h = H.host_only_canvas(256, 256)
    try:
        H._run(prod, h, script)
        n = prod.cv_path_edges(h, None, 0)
        e = np.zeros((max(n, 1), 4), np.float32)
        assert prod.cv_path_edges(h, e.ctypes.data, n) == n
        return e[:n]
    finally:
        prod.cv_destroy(h)


def _oracle_inside(edges, q):
    orc = H.oracle_library()
    out = np.zeros(len(q), np.uint8)
    orc.oracle_points_in_path(edges.ctypes.data, len(edges), q.ctypes.data, len(q), out.ctypes.data)
    return out


@pytest.mark.parametrize("kind", SCENES)
def test_oracle_hit_test_is_the_references(kind):
    ref = H.reference_library()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    prod = H.product_library()
    script = scene(kind)
    q = queries(6000)
    edges = _edges(prod, script)
    got = _oracle_inside(edges, q)
    h = _path_on(ref, script, lambda: ref.cv_create(256, 256))
    try:
        want = np.zeros(len(q), np.uint8)
        assert ref.cv_points_in_path(h, q.ctypes.data, len(q), want.ctypes.data) == 0
    finally:
        ref.cv_destroy(h)
    assert np.array_equal(got, want)
    if kind not in ("empty", "degenerate"):
        assert 0 < got.sum() < len(q)
    # the product's own single-point query (host) agrees as well
    h = _path_on(prod, script, lambda: H.host_only_canvas(256, 256))
    try:
        for i in range(0, len(q), 97):
            assert prod.cv_is_point_in_path(h, float(q[i, 0]), float(q[i, 1])) == int(want[i])
    finally:
        prod.cv_destroy(h)


def test_exact_edge_and_vertex_rules_are_exercised():
    """The lattice really lands on edges: some answers flip if `on an edge` were not counted."""
    prod = H.product_library()
    edges = _edges(prod, scene("rects_and_hole"))
    q = np.array([[32, 32], [32, 100], [192, 160], [64, 64], [96, 64], [100, 100], [31.999, 50], [192.001, 50]], np.float32)
    assert _oracle_inside(edges, q).tolist() == [1, 1, 1, 1, 1, 1, 0, 0]


def test_bulk_query_without_a_device_fails_loudly():
    prod = H.product_library()
    h = H.host_only_canvas(64, 64)
    try:
        q = np.zeros((4, 2), np.float32)
        out = np.zeros(4, np.uint8)
        assert prod.cv_points_in_path(h, q.ctypes.data, 4, out.ctypes.data) == -1      # CB200_ERR_NO_DEVICE
    finally:
        prod.cv_destroy(h)


# ------------------------------------------------------------------------ GPU ----

@pytest.fixture(scope="module")
def lib():
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    return lib


@pytest.mark.gpu
@pytest.mark.parametrize("kind", SCENES)
def test_gpu_hit_test_equals_oracle(lib, kind):
    script = scene(kind)
    q = queries(20000 if kind == "long_path" else 200000, seed=5)     # the CPU oracle is O(points x edges)
    want = _oracle_inside(_edges(lib, script), q)
    h = _path_on(lib, script, lambda: lib.cv_create(256, 256))
    try:
        got = np.full(len(q), 7, np.uint8)
        assert lib.cv_points_in_path(h, q.ctypes.data, len(q), got.ctypes.data) == 0, lib.cv_last_error()
        assert np.array_equal(got, want)
        few = np.full(50, 7, np.uint8)                       # few queries: the edges are split across CTAs instead
        assert lib.cv_points_in_path(h, q.ctypes.data, 50, few.ctypes.data) == 0
        assert np.array_equal(few, want[:50])
        assert lib.cv_points_in_path(h, q.ctypes.data, 0, few.ctypes.data) == 0
    finally:
        lib.cv_destroy(h)


@pytest.mark.gpu
def test_gpu_hit_test_through_the_c_abi_with_timing(lib):
    edges = _edges(lib, scene("long_path"))
    q = queries(500000, seed=9)
    want = _oracle_inside(edges, q[:20000])
    dev = C.c_void_p()
    assert lib.cb200_canvas_create(64, 64, 0, C.byref(dev)) == 0
    try:
        got = np.zeros(len(q), np.uint8)
        ms = C.c_float(0)
        assert lib.cb200_hit_test(dev, edges.ctypes.data, len(edges), q.ctypes.data, len(q), got.ctypes.data, C.byref(ms)) == 0
        assert np.array_equal(got[:20000], want)
        assert ms.value > 0.0
        again = np.zeros(len(q), np.uint8)                    # deterministic
        assert lib.cb200_hit_test(dev, edges.ctypes.data, len(edges), q.ctypes.data, len(q), again.ctypes.data, None) == 0
        assert np.array_equal(got, again)
    finally:
        lib.cb200_canvas_destroy(dev)


@pytest.mark.gpu
def test_python_mirror_bulk_query(lib):
    import canvas_ity_b200 as cb
    c = cb.Canvas(256, 256)
    try:
        c.move_to(30, 30); c.line_to(220, 40); c.line_to(128, 230); c.close_path()
        q = queries(5000, seed=11)
        many = c.points_in_path(q)
        assert many.dtype == bool and 0 < many.sum() < len(q)
        for i in range(0, len(q), 211):                       # the single-point host query says the same
            assert c.is_point_in_path(float(q[i, 0]), float(q[i, 1])) == bool(many[i])
    finally:
        c.close()
