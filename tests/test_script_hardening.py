"""Malformed canvas scripts end the replay with -1 before anything is dispatched (include/canvas_b200_api.h:
cv_run_script), image records whose blob is shorter than their geometry included, and the Python mirror refuses
short buffers instead of forwarding them (round-1 advisor findings on csrc/host/script.hpp and canvas.py)."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from canvas_ity_b200.script import ScriptWriter


def _run(script):
    lib = H.product_library()
    h = H.host_only_canvas(64, 64)
    try:
        return lib.cv_run_script(h, script, len(script), None, 0, None)
    finally:
        lib.cv_destroy(h)


def _image_op(name, w, h, stride, n_bytes):
    s = ScriptWriter()
    if name == "SET_PATTERN":
        s.ints(name, 0, w, h, stride, 0)
    elif name == "DRAW_IMAGE":
        s.ints(name, w, h, stride); s.raw("4f", 0, 0, 10, 10)
    else:
        s.ints(name, w, h, stride, 0, 0)
    s.blob(bytes(n_bytes))
    return s.take()


@pytest.mark.parametrize("name", ["SET_PATTERN", "DRAW_IMAGE", "PUT_IMAGE_DATA"])
def test_image_blob_must_cover_its_geometry(name):
    assert _run(_image_op(name, 4, 4, 16, 64)) == 1                    # exact
    assert _run(_image_op(name, 4, 4, 20, 3 * 20 + 16)) == 1           # padded rows: (h - 1) * stride + 4 w
    assert _run(_image_op(name, 4, 4, 16, 63)) == -1                   # one byte short
    assert _run(_image_op(name, 4, 4, 1 << 20, 64)) == -1              # stride far beyond the blob
    assert _run(_image_op(name, 4, 4, -16, 64)) == -1                  # a negative pitch cannot be carried
    assert _run(_image_op(name, 0, 4, 16, 0)) == 1                     # no-op geometry reads nothing (hpp:2847)
    assert _run(_image_op(name, 4, 4, 16, 0)) == 1                     # null image: the reference's silent no-op


def test_get_image_data_inside_a_script_sizes_its_scratch_by_geometry():
    s = ScriptWriter()
    s.ints("GET_IMAGE_DATA", 8, 8, 4, 0, 0); s.raw("I", 0)             # stride < 4 * width: rows overlap
    assert _run(s.take()) == 1
    s = ScriptWriter()
    s.ints("GET_IMAGE_DATA", 8, 8, -32, 0, 0); s.raw("I", 0)
    assert _run(s.take()) == -1


def test_truncated_records_do_not_dispatch():
    s = ScriptWriter()
    s.ints("SET_LINE_DASH", 1000)                                        # claims 1000 segments, carries 2
    s.raw("2f", 1.0, 2.0)
    assert _run(s.take()) == -1
    full = H.tiger_script(64, 64)
    for cut in (1, 3, 7):                                                # inside the first record (TRANSLATE: 1 + 8 bytes)
        assert _run(full[:cut]) == -1
    assert _run(full[:9]) == 1                                           # exactly one whole record
    for cut in range(10, 4000, 37):                                      # anywhere: an error or a whole prefix, never a crash
        assert _run(full[:cut]) >= -1
    assert _run(full) > 300


def test_python_mirror_rejects_short_image_buffers():
    import canvas_ity_b200 as cb
    from canvas_ity_b200.canvas import _image_bytes
    img = np.zeros((4, 4, 4), np.uint8)
    assert len(_image_bytes(img, 4, 4, 16)) == 64
    with pytest.raises(ValueError):
        _image_bytes(img, 4, 5, 16)
    with pytest.raises(ValueError):
        _image_bytes(img, 4, 4, -16)
    assert _image_bytes(img, 0, 4, 16) is not None                       # no-op geometry
    assert cb is not None
