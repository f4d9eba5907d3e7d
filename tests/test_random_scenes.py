"""Randomised parity: seeded scenes that mix what the 76 reference tests exercise one at a time -- transforms,
all eleven composite operations, solid / linear / radial / pattern brushes, dashes, joins and caps, clip paths,
shadows, global alpha, text -- through the CUDA back end and through the oracle on the same lowered frames
(float framebuffer within 1e-4 relative, RGBA8 within 1 LSB).  The oracle side of the same scenes is pinned to
the reference build on the CPU (`test_oracle_matches_reference_on_random_scenes`)."""
import numpy as np
import pytest

from tests import harness as H

SIZE = 192
OPS = [1, 2, 3, 4, 7, 10, 11, 12, 13, 14, 15]


def random_scene(seed):
    rng = np.random.default_rng(1000 + seed)
    u = lambda lo=0.0, hi=1.0: float(rng.uniform(lo, hi))
    w = H.ScriptWriter()
    # a background so that destination-dependent operations have something to work on
    w.ints("SET_COLOR", 0); w.raw("4f", u(), u(), u(), u(0.3, 1.0))
    w.floats("FILL_RECTANGLE", u(-20, 40), u(-20, 40), u(80, 220), u(80, 220))
    image = rng.integers(0, 256, (8, 8, 4), dtype=np.uint8)
    for _ in range(int(rng.integers(3, 7))):
        w.bare("SAVE")
        if rng.random() < 0.6:
            w.floats("TRANSLATE", u(-20, 60), u(-20, 60)); w.floats("ROTATE", u(-0.8, 0.8)); w.floats("SCALE", u(0.5, 1.8), u(0.5, 1.8))
        if rng.random() < 0.3:                                # clip to a blob
            w.bare("BEGIN_PATH"); w.floats("ARC", u(40, 150), u(40, 150), u(30, 90), 0.0, 6.2831855, 0); w.bare("CLIP")
        w.ints("SET_COMPOSITE", int(rng.choice(OPS)))
        w.floats("SET_GLOBAL_ALPHA", u(0.3, 1.0))
        if rng.random() < 0.25:
            w.floats("SET_SHADOW_COLOR", u(), u(), u(), u(0.4, 1.0)); w.floats("SET_SHADOW_BLUR", u(0.0, 9.0))
            w.floats("SET_SHADOW_OFFSET_X", u(-6, 6)); w.floats("SET_SHADOW_OFFSET_Y", u(-6, 6))
        stroke = rng.random() < 0.5
        which = 1 if stroke else 0
        kind = rng.integers(0, 4)
        if kind == 0:
            w.ints("SET_COLOR", which); w.raw("4f", u(), u(), u(), u(0.2, 1.0))
        elif kind == 1:
            w.ints("SET_LINEAR_GRADIENT", which); w.raw("4f", u(0, SIZE), u(0, SIZE), u(0, SIZE), u(0, SIZE))
        elif kind == 2:
            w.ints("SET_RADIAL_GRADIENT", which); w.raw("6f", u(40, 150), u(40, 150), u(0, 20), u(40, 150), u(40, 150), u(30, 120))
        else:
            w.ints("SET_PATTERN", which, 8, 8, 32, int(rng.integers(0, 4))); w.blob(image.tobytes())
        if kind in (1, 2):
            for o in sorted(rng.uniform(0, 1, int(rng.integers(2, 5)))):
                w.ints("ADD_COLOR_STOP", which); w.raw("5f", float(o), u(), u(), u(), u(0.2, 1.0))
        w.bare("BEGIN_PATH")
        for _ in range(int(rng.integers(1, 3))):
            w.floats("MOVE_TO", u(0, SIZE), u(0, SIZE))
            for _ in range(int(rng.integers(2, 6))):
                r = rng.random()
                if r < 0.35:
                    w.floats("LINE_TO", u(0, SIZE), u(0, SIZE))
                elif r < 0.7:
                    w.floats("BEZIER_CURVE_TO", *[u(-20, SIZE + 20) for _ in range(6)])
                elif r < 0.85:
                    w.floats("QUADRATIC_CURVE_TO", *[u(0, SIZE) for _ in range(4)])
                else:
                    w.floats("ARC", u(30, 160), u(30, 160), u(5, 60), u(0, 6.28), u(0, 6.28), int(rng.integers(0, 2)))
            if rng.random() < 0.5:
                w.bare("CLOSE_PATH")
        if stroke:
            w.floats("SET_LINE_WIDTH", u(0.5, 14.0)); w.ints("SET_LINE_JOIN", int(rng.integers(0, 3))); w.ints("SET_LINE_CAP", int(rng.integers(0, 3)))
            w.floats("SET_MITER_LIMIT", u(1.0, 12.0))
            if rng.random() < 0.4:
                dashes = [u(2, 20) for _ in range(int(rng.integers(1, 5)))]
                w.ints("SET_LINE_DASH", len(dashes)); w.raw("%df" % len(dashes), *dashes)
                w.floats("SET_LINE_DASH_OFFSET", u(0, 30))
            w.bare("STROKE")
        else:
            w.bare("FILL")
        w.bare("RESTORE")
    if rng.random() < 0.5:
        w.floats("SET_FONT", u(14, 40)); w.raw("B", 1); w.blob(H.font_a())
        w.ints("SET_COLOR", 0); w.raw("4f", u(), u(), u(), 1.0)
        w.floats("FILL_TEXT", u(0, 80), u(30, 170), 1.0e30); w.blob(b"CDE nst*")
    return w.take()


def random_scene_wide(seed):
    """A second generator: odd canvas sizes (tiles cut at the right / bottom edge), geometry far outside the
    canvas, degenerate segments, rectangles, arc_to, nested clips, draw_image / put_image_data / clear_rectangle,
    stroked and aligned text with a maximum width.  Returns (script, width, height)."""
    rng = np.random.default_rng(5000 + seed)
    u = lambda lo=0.0, hi=1.0: float(rng.uniform(lo, hi))
    W, Hh = int(rng.integers(33, 300)), int(rng.integers(33, 300))
    w = H.ScriptWriter()
    image = rng.integers(0, 256, (6, 5, 4), dtype=np.uint8)
    if rng.random() < 0.7:
        w.ints("SET_COLOR", 0); w.raw("4f", u(), u(), u(), u(0.2, 1.0)); w.floats("FILL_RECTANGLE", u(-30, 30), u(-30, 30), u(40, 330), u(40, 330))
    depth = 0
    for _ in range(int(rng.integers(4, 10))):
        r = rng.random()
        if r < 0.15 and depth < 3:
            w.bare("SAVE"); depth += 1
            w.floats("TRANSLATE", u(-40, 80), u(-40, 80)); w.floats("ROTATE", u(-3.2, 3.2)); w.floats("SCALE", u(0.3, 2.5), u(0.3, 2.5))
            continue
        if r < 0.25 and depth > 0:
            w.bare("RESTORE"); depth -= 1
            continue
        if r < 0.33:
            w.bare("BEGIN_PATH"); w.floats("RECTANGLE", u(-50, W), u(-50, Hh), u(10, 250), u(10, 250))
            if rng.random() < 0.5:
                w.floats("ARC", u(0, W), u(0, Hh), u(10, 120), 0.0, 6.2831855, 0)
            w.bare("CLIP")
            continue
        w.ints("SET_COMPOSITE", int(rng.choice(OPS)))
        w.floats("SET_GLOBAL_ALPHA", u(0.2, 1.0))
        if rng.random() < 0.3:
            w.floats("SET_SHADOW_COLOR", u(), u(), u(), u(0.3, 1.0)); w.floats("SET_SHADOW_BLUR", u(0.0, 14.0))
            w.floats("SET_SHADOW_OFFSET_X", u(-15, 15)); w.floats("SET_SHADOW_OFFSET_Y", u(-15, 15))
        else:
            w.floats("SET_SHADOW_COLOR", 0.0, 0.0, 0.0, 0.0)
        which = int(rng.integers(0, 2))
        w.ints("SET_COLOR", which); w.raw("4f", u(), u(), u(), u(0.2, 1.0))
        if rng.random() < 0.3:
            w.ints("SET_LINEAR_GRADIENT", which); w.raw("4f", u(-50, W + 50), u(-50, Hh + 50), u(-50, W + 50), u(-50, Hh + 50))
            for o in sorted(rng.uniform(0, 1, int(rng.integers(1, 4)))):
                w.ints("ADD_COLOR_STOP", which); w.raw("5f", float(o), u(), u(), u(), u(0.2, 1.0))
        kind = rng.random()
        w.floats("SET_LINE_WIDTH", float(rng.choice([0.3, 1.0, 2.5, 9.0, 31.0]))); w.ints("SET_LINE_JOIN", int(rng.integers(0, 3)))
        w.ints("SET_LINE_CAP", int(rng.integers(0, 3))); w.floats("SET_MITER_LIMIT", u(1.0, 20.0))
        if rng.random() < 0.3:
            dashes = [u(0.5, 25) for _ in range(int(rng.integers(1, 6)))]
            w.ints("SET_LINE_DASH", len(dashes)); w.raw("%df" % len(dashes), *dashes); w.floats("SET_LINE_DASH_OFFSET", u(-20, 40))
        else:
            w.ints("SET_LINE_DASH", 0)
        if kind < 0.12:
            w.floats("FILL_RECTANGLE" if which == 0 else "STROKE_RECTANGLE", u(-60, W), u(-60, Hh), u(-80, 260), u(-80, 260))
        elif kind < 0.2:
            w.floats("CLEAR_RECTANGLE", u(-20, W), u(-20, Hh), u(5, 120), u(5, 120))
        elif kind < 0.3:
            w.ints("DRAW_IMAGE", 5, 6, 20); w.raw("4f", u(-30, W), u(-30, Hh), u(-120, 200), u(-120, 200)); w.blob(image.tobytes())
        elif kind < 0.36:
            w.ints("PUT_IMAGE_DATA", 5, 6, 20, int(rng.integers(-4, W)), int(rng.integers(-4, Hh))); w.blob(image.tobytes())
        elif kind < 0.5:
            w.floats("SET_FONT", u(8, 60)); w.raw("B", 1); w.blob(H.font_a())
            w.ints("SET_TEXT_ALIGN", int(rng.integers(0, 3))); w.ints("SET_TEXT_BASELINE", int(rng.integers(0, 5)))
            w.floats("FILL_TEXT" if which == 0 else "STROKE_TEXT", u(-20, W), u(0, Hh), float(rng.choice([1.0e30, 40.0, 150.0]))); w.blob(b"CDEF GHI*nst")
        else:
            w.bare("BEGIN_PATH")
            for _ in range(int(rng.integers(1, 4))):
                big = 2000.0 if rng.random() < 0.15 else 0.0            # now and then far outside the canvas
                px, py = u(-40 - big, W + 40 + big), u(-40 - big, Hh + 40 + big)
                w.floats("MOVE_TO", px, py)
                for _ in range(int(rng.integers(1, 7))):
                    t = rng.random()
                    if t < 0.3:
                        w.floats("LINE_TO", u(-40 - big, W + 40 + big), u(-40 - big, Hh + 40 + big))
                    elif t < 0.4:
                        w.floats("LINE_TO", px, py)                      # back to the start: zero-length / doubled segments
                    elif t < 0.65:
                        w.floats("BEZIER_CURVE_TO", *[u(-60 - big, W + 60 + big) for _ in range(6)])
                    elif t < 0.8:
                        w.floats("ARC_TO", u(0, W), u(0, Hh), u(0, W), u(0, Hh), u(0, 60))
                    else:
                        w.floats("ARC", u(0, W), u(0, Hh), u(0, 80), u(-7, 7), u(-7, 7), int(rng.integers(0, 2)))
                if rng.random() < 0.5:
                    w.bare("CLOSE_PATH")
            w.bare("FILL" if which == 0 else "STROKE")
    return w.take(), W, Hh


# Seeds 51, 672, 679 and 695 of the wide generator were open defects of the CUDA path in round 1 (shadow working
# rectangle; per-edge clip arithmetic): they sit inside the committed range on purpose.
SEEDS = list(range(300))
WIDE_SEEDS = list(range(900))
INTEGER_SEEDS = list(range(300))
CHUNK = 50


def _scene(generator, seed):
    if generator == "plain":
        return random_scene(seed), SIZE, SIZE
    if generator == "wide":
        return random_scene_wide(seed)
    return integer_scene(seed)


def _chunks(generator, seeds):
    return [(generator, seeds[i:i + CHUNK]) for i in range(0, len(seeds), CHUNK)]


ALL_CHUNKS = _chunks("plain", SEEDS) + _chunks("wide", WIDE_SEEDS) + _chunks("integer", INTEGER_SEEDS)
CHUNK_IDS = ["%s-%d" % (g, c[0]) for g, c in ALL_CHUNKS]


@pytest.mark.parametrize("generator,seeds", ALL_CHUNKS, ids=CHUNK_IDS)
def test_oracle_matches_reference_on_random_scenes(generator, seeds):
    ref = H.reference_library()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    failed = []
    for seed in seeds:
        script, w, h = _scene(generator, seed)
        want = H.render_script(ref, script, w, h)
        got = H.render_oracle(script, w, h)
        nbad, worst = H.float_mismatch(got["f32"], want["f32"])
        n8 = H.rgba8_mismatch(got["rgba8"], want["rgba8"])[2]
        if nbad or n8:
            failed.append((seed, nbad, worst, n8))
    assert not failed, failed


@pytest.mark.gpu
@pytest.mark.parametrize("generator,seeds", ALL_CHUNKS, ids=CHUNK_IDS)
def test_gpu_matches_oracle_on_random_scenes(generator, seeds):
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    failed = []
    for seed in seeds:
        script, w, h = _scene(generator, seed)
        got = H.render_script(lib, script, w, h)
        want = H.render_oracle(script, w, h)
        nbad, worst = H.float_mismatch(got["f32"], want["f32"])
        n8 = H.rgba8_mismatch(got["rgba8"], want["rgba8"])[2]
        if nbad or n8:
            failed.append((seed, nbad, worst, n8))
    assert not failed, "(seed, floats off, max |diff|, pixels beyond 1 LSB): %s" % failed


@pytest.mark.gpu
@pytest.mark.parametrize("generator,seeds", _chunks("wide", [51, 672, 679, 695]) + _chunks("integer", list(range(300, 340))), ids=["round1-open-seeds", "integer-300"])
def test_gpu_matches_reference_on_formerly_open_seeds(generator, seeds):
    """The four seeds DESIGN.md listed as open after round 1, and a further block of integer scenes, against the
    REFERENCE build itself (not the oracle): float framebuffer within 1e-4 relative, RGBA8 within 1 LSB."""
    lib, ref = H.product_library(), H.reference_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    failed = []
    for seed in seeds:
        script, w, h = _scene(generator, seed)
        got = H.render_script(lib, script, w, h)
        want = H.render_script(ref, script, w, h)
        nbad, worst = H.float_mismatch(got["f32"], want["f32"])
        n8 = H.rgba8_mismatch(got["rgba8"], want["rgba8"])[2]
        if nbad or n8:
            failed.append((seed, nbad, worst, n8))
    assert not failed, "(seed, floats off, max |diff|, pixels beyond 1 LSB): %s" % failed


def integer_scene(seed):
    """Shadowed polygons and rectangles on INTEGER coordinates with integer shadow offsets, partly outside the
    canvas: crossings and vertices land exactly on pixel and canvas boundaries, where a working rectangle that is
    one row or column off shows.  Returns (script, width, height)."""
    rng = np.random.default_rng(9000 + seed)
    W, Hh = int(rng.integers(40, 200)), int(rng.integers(40, 200))
    w = H.ScriptWriter()
    for _ in range(int(rng.integers(2, 6))):
        w.ints("SET_COMPOSITE", int(rng.choice([1, 2, 3, 4, 7, 14])))
        w.floats("SET_SHADOW_COLOR", 0.0, 0.0, 0.0, 1.0); w.floats("SET_SHADOW_BLUR", float(rng.choice([0.0, 2.0, 4.0, 7.0])))
        w.floats("SET_SHADOW_OFFSET_X", float(rng.integers(-12, 13))); w.floats("SET_SHADOW_OFFSET_Y", float(rng.integers(-12, 13)))
        w.ints("SET_COLOR", 0); w.raw("4f", 0.5, 0.2, 0.8, 1.0)
        if rng.random() < 0.5:
            w.floats("FILL_RECTANGLE", float(rng.integers(-60, W)), float(rng.integers(-60, Hh)), float(rng.integers(5, 2 * W)), float(rng.integers(5, 2 * Hh)))
        else:
            w.bare("BEGIN_PATH")
            w.floats("MOVE_TO", float(rng.integers(-50, W + 50)), float(rng.integers(-50, Hh + 50)))
            for _ in range(int(rng.integers(2, 7))):
                w.floats("LINE_TO", float(rng.integers(-50, W + 50)), float(rng.integers(-50, Hh + 50)))
            w.bare("CLOSE_PATH"); w.bare("FILL")
    return w.take(), W, Hh


def shadow_box_mismatches(scenes):
    """(mismatching draws, shadowed draws): the run bounding box render_shadow derives from the reference's polygon
    clip (oracle) against what the product's own clip / scanline / boundary-segment code enters for the same
    outlines (csrc/device/edge_clip.cuh built for the host: cb200_debug_shadow_box).  CPU only."""
    import ctypes as C
    import struct
    orc, prod = H.oracle_library(), H.product_library()
    orc.oracle_debug_shadow_boxes.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int * 8)]
    orc.oracle_debug_shadow_boxes.restype = C.c_int
    per_loop = C.cast(prod.cb200_debug_shadow_box, C.c_void_p)
    bad, checked = [], 0
    for script, w, h in scenes:
        for fr in H.lower_script(script, w, h):
            draws = bytes(fr.parts["draws"])
            for di in range(fr.n_draws):
                rec = draws[di * 140:(di + 1) * 140]
                kind = struct.unpack_from("<I", rec, 0)[0]
                color_a, off_x, off_y, blur = struct.unpack_from("<4f", rec, 120)
                if kind == 2 or color_a == 0.0 or (blur == 0.0 and off_x == 0.0 and off_y == 0.0):
                    continue
                out = (C.c_int * 8)()
                orc.oracle_debug_shadow_boxes(C.addressof(fr.frame), di, w, h, per_loop, C.byref(out))
                checked += 1
                if list(out[0:4]) != list(out[4:8]):
                    bad.append((w, h, di, list(out[0:4]), list(out[4:8])))
    return bad, checked


def test_shadow_working_rectangle_equals_the_reference_clip():
    """render_shadow's working rectangle (hpp:2409-2423) comes from the runs of the Sutherland-Hodgman-clipped
    outline (hpp:2208-2229); the CUDA rasteriser clips edge by edge and rebuilds the clip's boundary segments from
    the loop's crossings.  The host build of that device code must enter exactly the reference's box on every
    shadowed draw of the fuzz scenes and of the integer-coordinate scenes."""
    scenes = [(random_scene(s), SIZE, SIZE) for s in range(120)] + [random_scene_wide(s) for s in range(300)] + \
             [integer_scene(s) for s in range(300)]
    bad, checked = shadow_box_mismatches(scenes)
    assert checked > 1000
    assert not bad, bad[:5]
