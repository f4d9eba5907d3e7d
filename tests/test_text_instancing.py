"""Device-side text: glyph outlines are cached once per font and text draws upload glyph
INSTANCES (include/canvas_b200.h, cb200_glyph_inst) that the device expands (k_glyph_instances).

The reference re-walks the TrueType outline on every draw (add_glyph hpp:1533-1696, text_to_lines
hpp:1793-1846); the instanced path must give the same control points bit for bit, so every check
here is `array_equal`, not a tolerance: instanced vs host-lowered text through the oracle (CPU) and
through the CUDA back end (GPU).  That the host lowering itself matches the reference is pinned by
tests/test_oracle_pinning.py (which now runs through the instanced path by default)."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H
from canvas_ity_b200 import _native

TEXT_TESTS = ["text_align", "text_baseline", "font", "fill_text", "stroke_text", "measure_text", "example_button"]


def _size(name):
    t = [x for x in H.manifest()["tests"] if x["name"] == name][0]
    return t["width"], t["height"]


class Atlas(C.Structure):          # cb200_glyph_atlas
    _fields_ = [("id", C.c_uint64), ("outlines", C.c_void_p), ("n_outlines", C.c_uint32),
                ("segs", C.c_void_p), ("n_segs", C.c_uint32), ("points", C.c_void_p), ("n_points", C.c_uint32)]


def _atlases(frame):
    n = frame.frame.n_atlases
    return C.cast(frame.frame.atlases, C.POINTER(Atlas * n)).contents if n else []


def test_atlas_record_matches_the_header():
    assert C.sizeof(Atlas) == _native.SIZEOF_GLYPH_ATLAS


@pytest.mark.parametrize("name", TEXT_TESTS)
def test_instanced_text_equals_host_lowering_on_the_oracle(name):
    w, h = _size(name)
    script = H.golden_script(name)
    on, off = H.lower_script(script, w, h), H.lower_script(script, w, h, instanced_text=False)
    assert sum(f.n_glyphs for f in on) > 0 and sum(f.n_glyphs for f in off) == 0
    # the glyph outlines no longer travel with every draw
    assert sum(f.n_points for f in on) < sum(f.n_points for f in off)
    assert sum(f.upload_bytes for f in on) < sum(f.upload_bytes for f in off)
    a, b = H.render_oracle(script, w, h), H.render_oracle(script, w, h, instanced_text=False)
    assert np.array_equal(a["f32"], b["f32"])
    assert np.array_equal(a["rgba8"], b["rgba8"])


def test_config5_scene_text_is_instanced_and_identical():
    for index in (0, 7, 123):
        script = H.config5_script(index)
        frames = H.lower_script(script, 256, 256)
        assert sum(f.n_glyphs for f in frames) > 0
        a, b = H.render_oracle(script, 256, 256), H.render_oracle(script, 256, 256, instanced_text=False)
        assert np.array_equal(a["f32"], b["f32"])


def test_canvases_with_the_same_font_share_one_atlas():
    """The cache is keyed by font content: a second canvas (or a repeated set_font) reuses the atlas, so
    its device copy is uploaded once and only ever extended."""
    a = H.lower_script(H.config5_script(3), 256, 256)
    b = H.lower_script(H.config5_script(4), 256, 256)
    ids_a = {at.id for f in a for at in _atlases(f)}
    ids_b = {at.id for f in b for at in _atlases(f)}
    assert len(ids_a) == 1 and ids_a == ids_b
    # append-only: the later snapshot holds at least what the earlier one did
    first = [at for f in a for at in _atlases(f)][0]
    later = [at for f in b for at in _atlases(f)][0]
    assert later.n_outlines >= first.n_outlines and later.n_points >= first.n_points
    # a different font gets a different atlas
    other = H.lower_script(H.golden_script("font"), *_size("font"))
    assert {at.id for f in other for at in _atlases(f)} - ids_a


def test_outline_records_are_self_consistent():
    frames = H.lower_script(H.golden_script("fill_text"), *_size("fill_text"))

    class Outline(C.Structure):
        _fields_ = [(n, C.c_uint32) for n in ("first_point", "n_points", "first_seg", "n_segs", "n_contours", "out_points")]

    class Seg(C.Structure):
        _fields_ = [(n, C.c_uint16) for n in ("from_a", "from_b", "to_a", "to_b", "ctrl", "flags")] + [("out", C.c_uint32)]

    assert C.sizeof(Outline) == _native.SIZEOF_GLYPH_OUTLINE and C.sizeof(Seg) == _native.SIZEOF_GLYPH_SEG
    at = [x for f in frames for x in _atlases(f)][0]
    outlines = C.cast(at.outlines, C.POINTER(Outline * at.n_outlines)).contents
    segs = C.cast(at.segs, C.POINTER(Seg * at.n_segs)).contents
    for o in outlines:
        assert o.out_points == o.n_contours + 3 * o.n_segs
        slots, firsts = set(), 0
        for k in range(o.first_seg, o.first_seg + o.n_segs):
            s = segs[k]
            assert max(s.from_a, s.from_b, s.to_a, s.to_b, s.ctrl) < o.n_points
            if s.flags & 2:
                firsts += 1
                slots.add(s.out - 1)
            slots.update((s.out, s.out + 1, s.out + 2))
        assert firsts == o.n_contours
        assert slots == set(range(o.out_points))          # every output point is written exactly once


# ------------------------------------------------------------------------ GPU ----

@pytest.fixture(scope="module")
def lib():
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    return lib


@pytest.mark.gpu
@pytest.mark.parametrize("name", TEXT_TESTS)
def test_instanced_text_equals_host_lowering_on_the_gpu(lib, name):
    w, h = _size(name)
    script = H.golden_script(name)
    a = H.render_script(lib, script, w, h)
    b = H.render_script(lib, script, w, h, instanced_text=False)
    assert np.array_equal(a["f32"], b["f32"])
    assert np.array_equal(a["rgba8"], b["rgba8"])
    want = H.render_oracle(script, w, h)
    nbad, worst = H.float_mismatch(a["f32"], want["f32"])
    assert nbad == 0, "max |diff| %.3g" % worst


@pytest.mark.gpu
def test_instanced_text_in_a_batch_and_across_frames(lib):
    """Config-5 canvases as one batch: every member's glyph instances are rebased onto the shared instance
    region and the one shared atlas; the result equals the canvases rendered alone with host-lowered text."""
    n = 12
    batch = lib.cv_batch_create(n, 256, 256, 0)
    assert batch
    try:
        for i in range(n):
            H._run(lib, lib.cv_batch_canvas(batch, i), H.config5_script(100 + i))
        assert lib.cv_batch_flush(batch) == 0
        for i in range(n):
            got = np.zeros((256, 256, 4), np.float32)
            assert lib.cv_batch_read_f32(batch, i, got.ctypes.data) == 0
            solo = H.render_script(lib, H.config5_script(100 + i), 256, 256, instanced_text=False)
            assert np.array_equal(got, solo["f32"]), "canvas %d" % i
    finally:
        lib.cv_batch_destroy(batch)


@pytest.mark.gpu
def test_atlas_grows_between_frames_of_one_canvas(lib):
    """New glyphs appear after the device copy of the atlas was made: the mirror is extended (or re-created)
    and earlier outlines stay valid."""
    font = H.font_a()
    h = lib.cv_create(256, 256)
    o = lib.cv_create(256, 256)
    lib.cv_set_text_instancing(o, 0)
    try:
        for canvas in (h, o):
            for k, text in enumerate([b"CDE", b"nst", b"CDEFGHI anstvy*"]):
                w = H.ScriptWriter()
                w.floats("SET_FONT", 28.0 + 6 * k); w.raw("B", 1); w.blob(font)
                w.ints("SET_COLOR", 0); w.raw("4f", 0.1, 0.2 + 0.2 * k, 0.7, 1.0)
                w.floats("FILL_TEXT", 8.0, 60.0 + 70 * k, 1.0e30); w.blob(text)
                H._run(lib, canvas, w.take())
                assert lib.cv_flush(canvas) == 0            # one frame per line of text
        a, b = np.zeros((256, 256, 4), np.float32), np.zeros((256, 256, 4), np.float32)
        assert lib.cv_read_f32(h, a.ctypes.data) == 0 and lib.cv_read_f32(o, b.ctypes.data) == 0
        assert a.any() and np.array_equal(a, b)
    finally:
        lib.cv_destroy(h)
        lib.cv_destroy(o)
