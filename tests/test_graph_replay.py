"""cb200_frame_replay as one CUDA graph launch: the graph captures exactly the stream sequence (header
restore, every kernel with its programmatic-dependent-launch edge, header readback), so replays through
the graph and through plain stream launches must leave bit-identical framebuffers -- with and without the
folded-in clear (two graphs), across the compositor builds, and after the resident frame is replaced."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from canvas_ity_b200 import _native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    return lib


def _scene(kind, size):
    if kind == "tiger":
        return H.tiger_script(size, size)
    if kind == "tiger_shadow":
        return H.tiger_script(size, size, global_alpha=0.9, shadow_blur=6.0, shadow_color=(0, 0, 0, 0.5))
    if kind == "gradient":
        return H.config4_script("linear", 15, size) + H.config4_script("radial", 10, size)
    if kind == "text_and_clip":
        return H.golden_script("example_button")
    raise ValueError(kind)


def _replays(lib, frame, size, graph, pattern):
    cv = C.c_void_p()
    assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
    try:
        assert lib.cb200_set_graph_replay(cv, graph) == 0
        assert lib.cb200_set_stage_timing(cv, 0) == 0
        assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
        for clear in pattern:
            assert lib.cb200_frame_replay(cv, clear) == 0, lib.cb200_last_error()
        out = np.zeros((size, size, 4), np.float32)
        assert lib.cb200_read_f32(cv, out.ctypes.data) == 0
        st = _native.Stats()
        assert lib.cb200_get_stats(cv, C.byref(st)) == 0
        return out, int(st.graph_replays), int(st.kernel_launches)
    finally:
        lib.cb200_canvas_destroy(cv)


@pytest.mark.parametrize("kind", ["tiger", "tiger_shadow", "gradient", "text_and_clip"])
def test_graph_replay_equals_stream_replay(lib, kind):
    size = 256 if kind == "text_and_clip" else 320
    frames = H.lower_script(_scene(kind, size), size, size)
    assert len(frames) == 1
    pattern = [1, 1, 1, 0, 0, 1, 0]                 # cleared and accumulating replays: both graphs get used
    with_graph, n_graph, launches_a = _replays(lib, frames[0], size, 1, pattern)
    streams, n_stream, launches_b = _replays(lib, frames[0], size, 0, pattern)
    assert n_graph >= len(pattern) - 2 and n_stream == 0         # the first replay verifies the frame on the stream path
    assert launches_a == launches_b
    assert with_graph.any() and np.array_equal(with_graph, streams)


def test_graph_is_rebuilt_when_the_resident_frame_changes(lib):
    size = 256
    a = H.lower_script(H.tiger_script(size, size), size, size)[0]
    b = H.lower_script(_scene("gradient", size), size, size)[0]
    cv = C.c_void_p()
    assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
    try:
        assert lib.cb200_set_stage_timing(cv, 0) == 0
        out = {}
        for name, frame in (("a", a), ("b", b), ("a2", a)):
            assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
            for _ in range(4):
                assert lib.cb200_frame_replay(cv, 1) == 0
            got = np.zeros((size, size, 4), np.float32)
            assert lib.cb200_read_f32(cv, got.ctypes.data) == 0
            out[name] = got
        st = _native.Stats()
        assert lib.cb200_get_stats(cv, C.byref(st)) == 0
        assert st.graph_replays >= 6
        assert np.array_equal(out["a"], out["a2"]) and not np.array_equal(out["a"], out["b"])
        want = H.render_script(lib, H.tiger_script(size, size), size, size)["f32"]
        assert np.array_equal(out["a"], want)
    finally:
        lib.cb200_canvas_destroy(cv)
