#!/usr/bin/env python
"""Device-side text: what the glyph outline cache + instancing buys.

Workload: one 1024x1024 canvas, N fill_text calls of 16 characters each (font_a of the reference suite,
random size / position / colour).  For text instancing on and off (cv_set_text_instancing) it reports
  * host recording time per text draw (front end only, frames tapped -- runs without a GPU),
  * bytes uploaded per frame,
and, when a GPU is present, the wall time of record + submit + sync and the device frame time, and
checks that both framebuffers are identical."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from canvas_ity_b200 import _native

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
SIZE = 1024
lib = _native.load()


def scene(n):
    state = [12345]

    def u():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xffffffff
        return (state[0] >> 8) / float(1 << 24)

    alphabet = " *CDEFGHIanstvy"
    w = H.ScriptWriter()
    w.floats("SET_FONT", 24.0); w.raw("B", 1); w.blob(H.font_a())
    for _ in range(n):
        w.floats("SET_FONT_RESIZE", 10 + 30 * u())                   # same font, new size
        w.ints("SET_COLOR", 0); w.raw("4f", u(), u(), u(), 0.5 + 0.5 * u())
        text = "".join(alphabet[int(u() * len(alphabet))] for _ in range(16))
        w.floats("FILL_TEXT", SIZE * u() - 40, SIZE * u(), 1.0e30); w.blob(text.encode())
    return w.take()


script = scene(N)
out = {"workload": "%d fill_text calls x 16 glyphs on a %dx%d canvas" % (N, SIZE, SIZE)}
for mode in (1, 0):
    key = "instanced" if mode else "host_lowered"
    best = 1e9
    for rep in range(5):
        frames = []

        @_native.FRAME_FN
        def on_frame(user, frame):
            frames.append((frame.contents.n_points, frame.contents.n_glyphs))

        h = lib.cv_create_tapped(SIZE, SIZE, C.cast(on_frame, C.c_void_p), None, None, None)
        lib.cv_set_text_instancing(h, mode)
        t0 = time.perf_counter()
        lib.cv_run_script(h, script, len(script), None, 0, None)
        lib.cv_flush(h)
        best = min(best, time.perf_counter() - t0)
        lib.cv_destroy(h)
    owned = H.lower_script(script, SIZE, SIZE, instanced_text=bool(mode))
    out[key] = {"host_record_us_per_text_draw": best / N * 1e6, "upload_bytes": sum(f.upload_bytes for f in owned),
                "uploaded_points": sum(f.n_points for f in owned), "glyph_instances": sum(f.n_glyphs for f in owned)}

if lib.cb200_device_count() > 0:
    images = {}
    for mode in (1, 0):
        key = "instanced" if mode else "host_lowered"
        h = lib.cv_create(SIZE, SIZE)
        lib.cv_set_text_instancing(h, mode)
        dev = lib.cv_device(h)
        walls, frames_ms, geom = [], [], []
        for rep in range(6):
            lib.cb200_clear(dev)
            t0 = time.perf_counter()
            lib.cv_run_script(h, script, len(script), None, 0, None)
            lib.cv_flush(h)
            lib.cb200_sync(dev)
            walls.append(time.perf_counter() - t0)
            st = _native.Stats()
            lib.cb200_get_stats(dev, C.byref(st))
            frames_ms.append(st.last_frame_ms); geom.append(st.geometry_ms)
        f = np.zeros((SIZE, SIZE, 4), np.float32)
        lib.cv_read_f32(h, f.ctypes.data)
        images[key] = f
        out[key].update({"record_submit_sync_ms": min(walls[1:]) * 1e3, "device_frame_ms": min(frames_ms[1:]),
                         "geometry_ms": min(geom[1:])})
        lib.cv_destroy(h)
    out["framebuffers_identical"] = bool(np.array_equal(images["instanced"], images["host_lowered"]))
print(json.dumps(out, indent=1))
