"""Shared helpers for tests/, bench.py and __graft_entry__.smoke().

Loads the three renderers that speak the flat canvas API / lowered-frame ABI:
  * product   canvas_ity_b200/libcanvas_b200.so   (CUDA; also the lowering front end)
  * oracle    oracle/liboracle.so                 (CPU restatement, test infrastructure)
  * reference oracle/_ref/libcanvas_ref.so        (the unmodified reference, when built)
and provides the comparison rules of the parity tests.
"""
import ctypes as C
import json
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from canvas_ity_b200 import _native           # noqa: E402
from canvas_ity_b200.script import ScriptWriter, OP   # noqa: E402

FLOAT_TOL = 1.0e-4          # BASELINE.json north_star: float coverage/colour within 1e-4 relative


# ------------------------------------------------------------------ libraries ----

def product_library():
    return _native.load()


_oracle = None


def oracle_library():
    """CPU restatement (oracle/oracle_raster.cpp); built by __graft_entry__.build()."""
    global _oracle
    if _oracle is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run __graft_entry__.build()")
        lib = C.CDLL(path)
        lib.oracle_canvas_create.restype = C.c_void_p
        lib.oracle_canvas_create.argtypes = [C.c_int, C.c_int]
        lib.oracle_canvas_destroy.argtypes = [C.c_void_p]
        lib.oracle_submit.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_read_f32.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_read_rgba8.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
        lib.oracle_write_rgba8.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
        lib.oracle_read_mask.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.oracle_points_in_path.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.oracle_debug_edges.restype = C.c_long
        lib.oracle_debug_edges.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_long]
        _oracle = lib
    return _oracle


_reference = {}


def reference_library(fast=False):
    """The real reference behind the flat API, or None when oracle/_ref was not built."""
    key = "fast" if fast else "exact"
    if key not in _reference:
        name = "libcanvas_ref_fast.so" if fast else "libcanvas_ref.so"
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(path):
            _reference[key] = None
        else:
            names = [n for n in _native.SIGNATURES if n.startswith("cv_")]
            _reference[key] = _native.bind(C.CDLL(path), names)
    return _reference[key]


# -------------------------------------------------------------------- scripts ----

def manifest():
    return json.load(open(os.path.join(GOLD, "manifest.json")))


def golden_script(name):
    return open(os.path.join(GOLD, "scripts", name + ".cvs"), "rb").read()


_golden_images = None


def golden_rgba8(name):
    global _golden_images
    if _golden_images is None:
        _golden_images = np.load(os.path.join(GOLD, "reference_rgba8.npz"))
    return _golden_images[name]


def tiger_prefix(width, height, global_alpha=None, shadow_blur=None, shadow_color=None):
    """Fit the 733x757 tiger into width x height."""
    w = ScriptWriter()
    s = np.float32(min(np.float32(width) / np.float32(733.0), np.float32(height) / np.float32(757.0)))
    tx = np.float32(0.5) * (np.float32(width) - np.float32(733.0) * s)
    ty = np.float32(0.5) * (np.float32(height) - np.float32(757.0) * s)
    w.floats("TRANSLATE", float(tx), float(ty))
    w.floats("SCALE", float(s), float(s))
    if global_alpha is not None:
        w.floats("SET_GLOBAL_ALPHA", global_alpha)
    if shadow_blur is not None:
        w.floats("SET_SHADOW_BLUR", shadow_blur)
    if shadow_color is not None:
        w.floats("SET_SHADOW_COLOR", *shadow_color)
    return w.take()


def tiger_script(width, height, **kw):
    return tiger_prefix(width, height, **kw) + golden_script("tiger")


# ------------------------------------------------------------------- rendering ----

def _run(lib, handle, script):
    q = (C.c_uint32 * (4 * 8192))()
    nq = C.c_int(0)
    n = lib.cv_run_script(handle, script, len(script), q, 8192, C.byref(nq))
    if n < 0:
        raise RuntimeError("malformed canvas script")
    return [tuple(q[i * 4:i * 4 + 3]) for i in range(min(nq.value, 8192))]


def render_script(lib, script, width, height, want_f32=True, instanced_text=True):
    """Replay `script` through a flat-API library (product or reference)."""
    h = lib.cv_create(width, height)
    if not h:
        raise RuntimeError(lib.cv_last_error().decode())
    try:
        if not instanced_text:
            lib.cv_set_text_instancing(h, 0)
        queries = _run(lib, h, script)
        out = {"queries": queries}
        img = np.zeros((height, width, 4), np.uint8)
        lib.cv_get_image_data(h, img.ctypes.data, width, height, 4 * width, 0, 0)
        out["rgba8"] = img
        if want_f32:
            f = np.zeros((height, width, 4), np.float32)
            rc = lib.cv_read_f32(h, f.ctypes.data)
            if rc != 0:
                raise RuntimeError(lib.cv_last_error().decode())
            out["f32"] = f
        return out
    finally:
        lib.cv_destroy(h)


def render_oracle(script, width, height, instanced_text=True):
    """Front-end lowering (product library, no device touched) -> oracle."""
    prod, orc = product_library(), oracle_library()
    o = orc.oracle_canvas_create(width, height)
    addr = lambda f: C.cast(f, C.c_void_p)
    h = prod.cv_create_tapped(width, height, addr(orc.oracle_tap_frame), addr(orc.oracle_tap_read),
                              addr(orc.oracle_tap_write), o)
    try:
        if not instanced_text:
            prod.cv_set_text_instancing(h, 0)
        queries = _run(prod, h, script)
        prod.cv_flush(h)
        f = np.zeros((height, width, 4), np.float32)
        orc.oracle_read_f32(o, f.ctypes.data)
        img = np.zeros((height, width, 4), np.uint8)
        orc.oracle_read_rgba8(o, img.ctypes.data, width, height, 4 * width, 0, 0)
        return {"f32": f, "rgba8": img, "queries": queries}
    finally:
        prod.cv_destroy(h)
        orc.oracle_canvas_destroy(o)


def lower_script(script, width, height, instanced_text=True):
    """Run `script` through the front end only and return the lowered frames it would submit
    (deep copies), plus the upload size.  No device is touched."""
    prod = product_library()
    frames = []

    @_native.FRAME_FN
    def on_frame(user, frame):
        frames.append(_native.OwnedFrame(frame.contents))

    h = prod.cv_create_tapped(width, height, C.cast(on_frame, C.c_void_p), None, None, None)
    try:
        if not instanced_text:
            prod.cv_set_text_instancing(h, 0)
        _run(prod, h, script)
        prod.cv_flush(h)
    finally:
        prod.cv_destroy(h)
    return frames


@_native.FRAME_FN
def _discard_frame(user, frame):
    pass


def host_only_canvas(width, height):
    """A front-end canvas without a device (frames are dropped): path building and host queries only."""
    return product_library().cv_create_tapped(width, height, C.cast(_discard_frame, C.c_void_p), None, None, None)


# ------------------------------------------------------------------ comparison ----

def float_mismatch(got, want, tol=FLOAT_TOL):
    """Count of float components off by more than tol * max(1, |want|) and the max abs diff."""
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    lim = tol * np.maximum(1.0, np.abs(want.astype(np.float64)))
    return int((d > lim).sum()), float(d.max() if d.size else 0.0)


def rgba8_mismatch(got, want):
    """Alpha-aware 8-bit difference (the reference's own hash weights colour by alpha,
    test/test.cpp:2382-2388): returns (max |d alpha|, max |d(colour*alpha)|/255, count > 1 LSB)."""
    g, w = got.astype(np.int32), want.astype(np.int32)
    da = np.abs(g[..., 3] - w[..., 3])
    dc = np.abs(g[..., :3] * g[..., 3:4] - w[..., :3] * w[..., 3:4]) / 255.0
    worst = np.maximum(da, dc.max(axis=-1))
    return int(da.max()), float(dc.max()), int((worst > 1.0).sum())


_roll_cache = {}


def hash_image(image):
    """The reference harness' locality-sensitive image hash (restated from test/test.cpp:2365-2407)."""
    h, w, _ = image.shape
    img = image.astype(np.int64)
    cur = img.copy()
    down = np.roll(img, -1, axis=0).copy()
    right = np.roll(img, -1, axis=1).copy()
    for arr in (cur, down, right):
        arr[..., :3] *= arr[..., 3:4]
    thr = np.array([8 * 255, 8 * 255, 8 * 255, 8], np.int64)
    edges = ((cur - down > thr * 16) * 128 | (cur - down > thr) * 64 | (down - cur > thr * 16) * 32 |
             (down - cur > thr) * 16 | (cur - right > thr * 16) * 8 | (cur - right > thr) * 4 |
             (right - cur > thr * 16) * 2 | (right - cur > thr) * 1).astype(np.uint64).reshape(-1)
    n = edges.size
    # xorshift state sequence is data independent: generate it vectorised in chunks
    rolls = _roll_cache.get(n)
    if rolls is None:
        state = 0xffffffff
        rolls = np.empty(n, np.uint64)
        for i in range(n):
            state ^= (state & 0x7ffff) << 13
            state ^= state >> 17
            state ^= (state & 0x7ffffff) << 5
            state &= 0xffffffff
            rolls[i] = state >> 27
        _roll_cache[n] = rolls
    e = edges
    r = rolls
    rotated = np.where(r > 0, ((e & (np.uint64(0xffffffff) >> r)) << r) | (e >> (np.uint64(32) - np.where(r > 0, r, 1))), e)
    rotated &= np.uint64(0xffffffff)
    return int(np.bitwise_xor.reduce(rotated))


def hamming(a, b):
    return bin((a ^ b) & 0xffffffff).count("1")


# ---------------------------------------------------------------- config 5 scenes ----

_font_a = None


def font_a():
    """The reference suite's main test font (test.cpp:114-170), recovered from a captured script."""
    global _font_a
    if _font_a is None:
        blob = golden_script("text_align")
        at = blob.index(b"\x00\x01\x00\x00")            # TrueType version tag = start of the font blob
        n = struct.unpack_from("<I", blob, at - 4)[0]
        _font_a = blob[at:at + n]
    return _font_a


def config5_script(index, with_text=True):
    """BASELINE.json config 5 / SURVEY 8d: canvas `index` of the 16384-canvas batch (256x256):
    8 random draws of 1-3 subpaths with 3-8 cubics each, half filled half stroked, plus one
    fill_text of 6 characters; everything from the LCG seeded with 0x9E3779B9 * (index + 1)."""
    state = [(0x9E3779B9 * (index + 1)) & 0xffffffff]

    def u():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xffffffff
        return (state[0] >> 8) / float(1 << 24)

    w = ScriptWriter()
    for _ in range(8):
        w.bare("BEGIN_PATH")
        for _ in range(1 + int(u() * 3)):
            w.floats("MOVE_TO", -32 + 320 * u(), -32 + 320 * u())
            for _ in range(3 + int(u() * 6)):
                w.floats("BEZIER_CURVE_TO", *[-32 + 320 * u() for _ in range(6)])
        stroke = u() < 0.5
        w.floats("SET_LINE_WIDTH", 0.5 + 11.5 * u())
        w.ints("SET_LINE_JOIN", int(u() * 3))
        w.ints("SET_LINE_CAP", int(u() * 3))
        w.ints("SET_COLOR", 1 if stroke else 0)
        w.raw("4f", u(), u(), u(), u())
        w.bare("STROKE" if stroke else "FILL")
    if with_text:
        alphabet = " *CDEFGHIanstvy"
        text = "".join(alphabet[int(u() * len(alphabet))] for _ in range(6))
        w.floats("SET_FONT", 16 + 48 * u()); w.raw("B", 1); w.blob(font_a())
        w.ints("SET_COLOR", 0); w.raw("4f", u(), u(), u(), 1.0)
        x, y = 192 * u(), 192 * u()
        w.floats("FILL_TEXT", x, y, 1.0e30); w.blob(text.encode())
    return w.take()


# ------------------------------------------------------------------ config 4 ----

def lcg_image(n, seed=12345):
    """n x n RGBA8 image of SURVEY 8d config 4: bytes = high byte of s = s * 1664525 + 1013904223."""
    s, out = seed, bytearray(n * n * 4)
    for i in range(len(out)):
        s = (s * 1664525 + 1013904223) & 0xffffffff
        out[i] = s >> 24
    return bytes(out)


_lcg_cache = {}


def config4_script(kind, op, size, image_size=None):
    """Full-canvas fill of SURVEY 8d config 4: kind in solid / linear / radial / image under composite op."""
    w = ScriptWriter()
    W = float(size)
    w.ints("SET_COMPOSITE", op)
    if kind == "solid":
        w.ints("SET_COLOR", 0); w.raw("4f", 0.2, 0.5, 0.9, 0.7)
        w.floats("FILL_RECTANGLE", 0, 0, W, W)
    elif kind == "linear":
        w.ints("SET_LINEAR_GRADIENT", 0); w.raw("4f", 0.1 * W, 0.2 * W, 0.9 * W, 0.8 * W)
        for o, c in ((0.0, (1, 0, 0, 1)), (0.5, (0, 1, 0, 0.5)), (1.0, (0, 0, 1, 1))):
            w.ints("ADD_COLOR_STOP", 0); w.raw("5f", o, *c)
        w.floats("FILL_RECTANGLE", 0, 0, W, W)
    elif kind == "radial":
        w.ints("SET_RADIAL_GRADIENT", 0); w.raw("6f", 0.4 * W, 0.4 * W, 0.05 * W, 0.5 * W, 0.5 * W, 0.5 * W)
        for o, c in ((0.0, (1, 1, 0, 1)), (1.0, (0, 1, 1, 0.3))):
            w.ints("ADD_COLOR_STOP", 0); w.raw("5f", o, *c)
        w.floats("FILL_RECTANGLE", 0, 0, W, W)
    elif kind == "image":
        n = image_size or (256 if size <= 2048 else 1024)
        if n not in _lcg_cache:
            _lcg_cache[n] = lcg_image(n)
        w.ints("DRAW_IMAGE", n, n, 4 * n); w.raw("4f", 0, 0, W, W); w.blob(_lcg_cache[n])
    else:
        raise ValueError(kind)
    return w.take()
