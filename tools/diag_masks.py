#!/usr/bin/env python
"""GPU: clip-mask planes of a fuzz scene, device vs oracle (same lowered frames), slot by slot."""
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
frames = H.lower_script(s, W, Hh)
cv = C.c_void_p()
assert lib.cb200_canvas_create(W, Hh, 0, C.byref(cv)) == 0
o = orc.oracle_canvas_create(W, Hh)
for fi, fr in enumerate(frames):
    assert lib.cb200_submit(cv, C.byref(fr.frame)) == 0, lib.cb200_last_error()
    orc.oracle_submit(o, C.byref(fr.frame))
    got = np.zeros((Hh, W, 4), np.float32); want = np.zeros((Hh, W, 4), np.float32)
    assert lib.cb200_read_f32(cv, got.ctypes.data) == 0
    orc.oracle_read_f32(o, want.ctypes.data)
    print("frame", fi, "draws", fr.n_draws, "fb max |diff|", float(np.abs(got - want).max()))
    for slot in range(1, 12):
        g = np.zeros((Hh, W), np.float32); w_ = np.zeros((Hh, W), np.float32)
        if lib.cb200_read_mask(cv, slot, g.ctypes.data) != 0: continue
        if orc.oracle_read_mask(o, slot, w_.ctypes.data) != 0: print("  slot", slot, "missing in the oracle"); continue
        d = np.abs(g - w_)
        ys, xs = np.nonzero(d > 1e-5)
        print("  mask slot", slot, "max |diff|", float(d.max()), "cells off", len(ys), "rows", sorted(set(ys.tolist()))[:8], "x range", (xs.min(), xs.max()) if len(xs) else None)
        for y, x in list(zip(ys, xs))[:5]: print("     (x %d, y %d) device %.6f oracle %.6f" % (x, y, g[y, x], w_[y, x]))
