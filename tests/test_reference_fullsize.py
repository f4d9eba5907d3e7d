"""The CUDA back end against the REFERENCE ITSELF at the benchmark's own sizes (BASELINE.json configs 1/3/4/5):
oracle/_ref/libcanvas_ref.so is the unmodified reference header compiled where it lies (oracle/Makefile), driven
through the same flat API and the same call stream as the product.  No oracle restatement and no product lowering
sits between the two sides here, so a lowering bug that only shows at 4096^2 / 8192^2 cannot hide.

Tolerances (BASELINE.json north_star): linear premultiplied float framebuffer within 1e-4 * max(1, |ref|), RGBA8
within 1 LSB (alpha-aware, the reference's own hash weights colour by alpha, test.cpp:2382-2388)."""
import ctypes as C

import numpy as np
import pytest

from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    lib, ref = H.product_library(), H.reference_library()
    assert lib.cb200_device_count() > 0, "gpu tests need a CUDA device; the back end has no CPU path"
    if ref is None:
        pytest.skip("oracle/_ref/libcanvas_ref.so was not built (it is built wherever /root/reference exists)")
    return lib, ref


def float_mismatch_rows(got, want, tol=H.FLOAT_TOL, rows=256):
    """harness.float_mismatch in row blocks: an 8192^2 x 4 framebuffer would need ~30 GB of float64 temporaries."""
    nbad, worst = 0, 0.0
    for y in range(0, got.shape[0], rows):
        n, w = H.float_mismatch(got[y:y + rows], want[y:y + rows], tol)
        nbad += n
        worst = max(worst, w)
    return nbad, worst


def rgba8_mismatch_rows(got, want, rows=512):
    da = dc = n8 = 0
    for y in range(0, got.shape[0], rows):
        a, c, n = H.rgba8_mismatch(got[y:y + rows], want[y:y + rows])
        da, dc, n8 = max(da, a), max(dc, c), n8 + n
    return da, dc, n8


def compare(libs, script, size, rgba8=True):
    lib, ref = libs
    got = H.render_script(lib, script, size, size)
    want = H.render_script(ref, script, size, size)
    nbad, worst = float_mismatch_rows(got["f32"], want["f32"])
    assert nbad == 0, "%d floats beyond 1e-4 relative, max |diff| %.3g" % (nbad, worst)
    if rgba8:
        da, dc, n8 = rgba8_mismatch_rows(got["rgba8"], want["rgba8"])
        assert n8 == 0 and da <= 1, "RGBA8: %d pixels beyond 1 LSB (max alpha diff %d, colour %.2f)" % (n8, da, dc)
    assert float(np.abs(want["f32"][::16, ::16]).sum()) > 0.0     # the reference drew something
    return worst


def test_tiger_4096_against_the_reference(libs):
    """BASELINE config `tiger_4096` (the bench's own workload): 305 draws, 2380 cubics."""
    compare(libs, H.tiger_script(4096, 4096), 4096)


def test_config3_shadows_4096_against_the_reference(libs):
    """BASELINE config 3: global_alpha 0.9, shadow_blur 16, shadow colour (0,0,0,0.5) -- 305 shadow planes, blur
    radius 7; the reference needs ~8 s for this frame."""
    compare(libs, H.tiger_script(4096, 4096, global_alpha=0.9, shadow_blur=16.0, shadow_color=(0, 0, 0, 0.5)), 4096)


@pytest.mark.parametrize("kind,op", [("linear", 15), ("radial", 10), ("image", 2)])
def test_config4_fill_8192_against_the_reference(libs, kind, op):
    """BASELINE config 4 at its full 8192^2: a full-canvas gradient / bicubic draw_image fill (1024^2 LCG image)
    under exclusive_or / lighter / source_copy over a translucent background."""
    size = 8192
    from canvas_ity_b200.script import ScriptWriter
    bg = ScriptWriter()
    bg.ints("SET_COLOR", 0); bg.raw("4f", 0.9, 0.8, 0.1, 0.6); bg.floats("FILL_RECTANGLE", 0, 0, float(size), float(size))
    compare(libs, bg.take() + H.config4_script(kind, op, size), size, rgba8=False)


def test_config5_batch_of_256_against_the_reference(libs):
    """BASELINE config 5: 256 of the 16384 seeded 256x256 canvases (8 random fills / strokes + one fill_text each)
    rendered as ONE batch frame, every member against the reference's rendering of the same canvas."""
    lib, ref = libs
    n, size = 256, 256
    scripts = [H.config5_script(i) for i in range(n)]
    batch = lib.cv_batch_create(n, size, size, 0)
    assert batch, lib.cv_last_error()
    try:
        for i, s in enumerate(scripts):
            H._run(lib, lib.cv_batch_canvas(batch, i), s)
        assert lib.cv_batch_flush(batch) == 0, lib.cv_last_error()
        failed = []
        for i, s in enumerate(scripts):
            got = np.zeros((size, size, 4), np.float32)
            assert lib.cv_batch_read_f32(batch, i, got.ctypes.data) == 0
            img = np.zeros((size, size, 4), np.uint8)
            assert lib.cv_batch_get_image_data(batch, i, img.ctypes.data, size, size, 4 * size, 0, 0) == 0
            want = H.render_script(ref, s, size, size)
            nbad, worst = H.float_mismatch(got, want["f32"])
            n8 = H.rgba8_mismatch(img, want["rgba8"])[2]
            if nbad or n8:
                failed.append((i, nbad, worst, n8))
        assert not failed, "(canvas, floats off, max |diff|, pixels beyond 1 LSB): %s" % failed[:10]
    finally:
        lib.cv_batch_destroy(batch)
