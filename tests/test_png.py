"""PNG off the device: cb200_encode_png writes the file the reference's test driver writes after
get_image_data (write_png, test/test.cpp:2415-2507) -- stored-deflate rows, Adler-32, CRC-32 -- in one kernel.

Byte/integer work, so everything is byte-exact:
  png_restated() below (struct + zlib.crc32 / zlib.adler32, an independent checksum implementation)
    == the reference's own write_png (oracle/_ref, ref_write_png)                         [CPU, pins the format]
    == cb200_encode_png over the same pixels (the canvas' own get_image_data)            [-m gpu]
Sizes cover rows that start on and off a word boundary (the file offset of a row is 62 + y (6 + 4 w)),
partial 256-pixel segments, single pixels and a multi-megapixel canvas."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

from tests import harness as H


def png_restated(rgba):
    """write_png (test/test.cpp:2415-2507) restated: IHDR, sRGB, one IDAT of stored blocks (one per row), IEND."""
    h, w, _ = rgba.shape

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)

    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))
    row = 1 + 4 * w
    parts = [b"\x78\x01"]
    for y in range(h):
        parts.append(struct.pack("<BHH", int(y + 1 == h), row, row ^ 0xffff))
        parts.append(raw[y * row:(y + 1) * row])
    parts.append(struct.pack(">I", zlib.adler32(raw) & 0xffffffff))
    body = b"".join(parts)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            chunk(b"sRGB", b"\x00") + chunk(b"IDAT", body) + chunk(b"IEND", b""))


def _ref_png(rgba, tmp_path):
    ref = H.reference_library()
    if ref is None:
        return None
    fn = ref.ref_write_png                       # the reference driver's write_png (oracle/ref_png.cpp)
    fn.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    fn.restype = None
    path = str(tmp_path / "ref.png")
    img = np.ascontiguousarray(rgba)
    fn(path.encode(), img.ctypes.data, img.shape[1], img.shape[0])
    return open(path, "rb").read()


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (257, 31), (300, 301), (1024, 2)])
def test_restated_png_is_the_reference_drivers(tmp_path, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    mine = png_restated(img)
    assert len(mine) == 76 + h * (6 + 4 * w)
    ref = _ref_png(img, tmp_path)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    assert mine == ref
    # and it is a valid zlib stream of filter-0 rows
    at = mine.index(b"IDAT") + 4
    n = struct.unpack(">I", mine[at - 8:at - 4])[0]
    rows = zlib.decompress(mine[at:at + n])
    assert rows == b"".join(b"\x00" + img[y].tobytes() for y in range(h))


def test_png_without_a_device_fails_loudly(tmp_path):
    prod = H.product_library()
    h = H.host_only_canvas(8, 8)
    try:
        assert prod.cv_write_png(h, str(tmp_path / "x.png").encode()) == -1          # CB200_ERR_NO_DEVICE
        assert not (tmp_path / "x.png").exists()
    finally:
        prod.cv_destroy(h)


# ------------------------------------------------------------------------ GPU ----

@pytest.fixture(scope="module")
def lib():
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    return lib


def _scene(w, h):
    """Something with gradients of alpha and colour everywhere (so every byte matters)."""
    s = H.ScriptWriter()
    s.ints("SET_LINEAR_GRADIENT", 0); s.raw("4f", 0.0, 0.0, float(w), float(h))
    for o, c in ((0.0, (1, 0, 0, 1)), (0.4, (0, 1, 0, 0.3)), (1.0, (0, 0.2, 1, 0.9))):
        s.ints("ADD_COLOR_STOP", 0); s.raw("5f", o, *c)
    s.floats("FILL_RECTANGLE", 0, 0, float(w), float(h))
    s.ints("SET_COLOR", 0); s.raw("4f", 0.9, 0.8, 0.1, 0.6)
    s.floats("ARC", 0.5 * w, 0.5 * h, 0.4 * min(w, h), 0.0, 6.2831855, 0)
    s.bare("FILL")
    return s.take()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(1, 1), (5, 3), (256, 256), (255, 64), (733, 757), (1000, 7), (2051, 300), (16383, 4)])
def test_gpu_png_is_byte_identical(lib, tmp_path, w, h):
    cv = lib.cv_create(w, h)
    assert cv
    try:
        H._run(lib, cv, _scene(w, h))
        img = np.zeros((h, w, 4), np.uint8)
        lib.cv_get_image_data(cv, img.ctypes.data, w, h, 4 * w, 0, 0)
        assert img.any()
        n = C.c_size_t(0)
        assert lib.cb200_encode_png(lib.cv_device(cv), None, 0, C.byref(n)) == 0
        assert n.value == 76 + h * (6 + 4 * w)
        buf = np.zeros(n.value, np.uint8)
        assert lib.cb200_encode_png(lib.cv_device(cv), buf.ctypes.data, n.value, None) == 0, lib.cv_last_error()
        want = png_restated(img)
        got = buf.tobytes()
        assert got[:56] == want[:56]
        assert got[-20:] == want[-20:], "checksums / trailer differ"
        assert got == want
        # a second encode (tables cached, accumulators reset) and the file writer give the same bytes
        path = str(tmp_path / "out.png")
        assert lib.cv_write_png(cv, path.encode()) == 0
        assert open(path, "rb").read() == want
        ref = _ref_png(img, tmp_path)
        if ref is not None:
            assert got == ref
    finally:
        lib.cv_destroy(cv)


@pytest.mark.gpu
def test_gpu_png_of_the_tiger_4096(lib):
    """Full size: 64 MiB of pixels, 16 segments per row, every row shifted by its own power of x."""
    size = 4096
    cv = lib.cv_create(size, size)
    try:
        H._run(lib, cv, H.tiger_script(size, size))
        img = np.zeros((size, size, 4), np.uint8)
        lib.cv_get_image_data(cv, img.ctypes.data, size, size, 4 * size, 0, 0)
        n = 76 + size * (6 + 4 * size)
        buf = np.zeros(n, np.uint8)
        assert lib.cb200_encode_png(lib.cv_device(cv), buf.ctypes.data, n, None) == 0
        assert buf.tobytes() == png_restated(img)
    finally:
        lib.cv_destroy(cv)


@pytest.mark.gpu
def test_gpu_png_rejects_what_the_format_cannot_hold(lib):
    n = C.c_size_t(0)
    band = C.c_void_p()
    assert lib.cb200_canvas_create_band(64, 64, 16, 16, 0, C.byref(band)) == 0
    try:
        assert lib.cb200_encode_png(band, None, 0, C.byref(n)) == -2                  # CB200_ERR_BAD_ARG
    finally:
        lib.cb200_canvas_destroy(band)
    wide = C.c_void_p()
    assert lib.cb200_canvas_create(16384, 2, 0, C.byref(wide)) == 0
    try:
        assert lib.cb200_encode_png(wide, None, 0, C.byref(n)) == -2
    finally:
        lib.cb200_canvas_destroy(wide)


@pytest.mark.gpu
def test_python_mirror_write_png(lib, tmp_path):
    import canvas_ity_b200 as cb
    c = cb.Canvas(300, 200)
    try:
        c.set_color(cb.fill_style, 0.2, 0.6, 0.9, 0.8)
        c.arc(150, 100, 80, 0.0, 6.2831855); c.fill()
        path = tmp_path / "mirror.png"
        c.write_png(path)
        assert path.read_bytes() == png_restated(c.get_image_data())
    finally:
        c.close()
