#!/usr/bin/env python
"""GPU: mismatching pixels (CUDA vs oracle) of a fuzz seed -> printed + gpurun_out/pixels_<seed>.npz, and a bit-for-bit
comparison of the device's outline edges with the oracle's for every draw of the scene (single-frame scenes)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tests.test_random_scenes import random_scene, random_scene_wide, SIZE

wide = '--wide' in sys.argv
seed = int([a for a in sys.argv[1:] if not a.startswith('--')][0])
lib, orc = H.product_library(), H.oracle_library()
if wide: s, W, Hh = random_scene_wide(seed)
else: s, W, Hh = random_scene(seed), SIZE, SIZE
got, want = H.render_script(lib, s, W, Hh)["f32"], H.render_oracle(s, W, Hh)["f32"]
d = np.abs(got.astype(np.float64) - want.astype(np.float64))
bad = (d > H.FLOAT_TOL * np.maximum(1.0, np.abs(want))).any(axis=-1)
ys, xs = np.nonzero(bad)
print("canvas", W, Hh, "pixels off", len(ys))
for y, x in list(zip(ys, xs))[:40]:
    print("  (x %d, y %d) got %s want %s" % (x, y, got[y, x], want[y, x]))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/pixels_%d.npz" % seed, ys=ys, xs=xs, got=got, want=want)
frames = H.lower_script(s, W, Hh)
print("frames", len(frames))
for fi, fr in enumerate(frames):
    cv = C.c_void_p()
    assert lib.cb200_canvas_create(W, Hh, 0, C.byref(cv)) == 0
    rc = lib.cb200_submit(cv, C.byref(fr.frame))
    if rc != 0:
        print("frame", fi, "submit failed (clip planes of an earlier frame?)", lib.cb200_last_error()); lib.cb200_canvas_destroy(cv); continue
    gotl = np.zeros((400000, 4), np.float32); jobs = np.zeros(400000, np.uint32)
    m = lib.cb200_debug_lines(cv, gotl.ctypes.data, jobs.ctypes.data, len(gotl))
    lib.cb200_canvas_destroy(cv)
    b = set(map(bytes, gotl[:m].view(np.uint8).reshape(m, 16)))
    for di in range(fr.n_draws):
        wantl = np.zeros((400000, 4), np.float32)
        n = orc.oracle_debug_edges(C.addressof(fr.frame), di, wantl.ctypes.data, len(wantl))
        wl = wantl[:n]
        inside = (wl.min(axis=1) >= 0) & (wl[:, [0, 2]].max(axis=1) <= W) & (wl[:, [1, 3]].max(axis=1) <= Hh) & (np.abs(wl[:, 3] - wl[:, 1]) >= 2.0e-5)
        a = set(map(bytes, wl[inside].view(np.uint8).reshape(int(inside.sum()), 16)))
        missing = a - b
        print("frame %d draw %d: %d oracle edges inside the canvas, %d not among the device's pieces" % (fi, di, len(a), len(missing)))
        for e in sorted(missing)[:6]: print("     oracle:", np.frombuffer(e, np.float32))
