#!/usr/bin/env python
"""GPU: where does the CUDA back end differ from the reference build on a full-size tiger frame?  Saves the
mismatching pixels (coordinates, got, want) under gpurun_out/ and bisects the call stream for the first draw
after which the two disagree.  usage: diag_fullsize.py [size] [shadow]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tools.diag_scene_lib import split, show, DRAWS

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
shadow = len(sys.argv) > 2 and sys.argv[2] == "shadow"
kw = dict(global_alpha=0.9, shadow_blur=16.0, shadow_color=(0, 0, 0, 0.5)) if shadow else {}
script = H.tiger_script(size, size, **kw)
lib, ref = H.product_library(), H.reference_library()


def mismatch(sc, keep=False):
    got = H.render_script(lib, sc, size, size)["f32"]
    want = H.render_script(ref, sc, size, size)["f32"]
    bad = np.zeros(got.shape[:2], bool)
    worst = 0.0
    for y in range(0, size, 256):
        d = np.abs(got[y:y + 256].astype(np.float64) - want[y:y + 256].astype(np.float64))
        lim = H.FLOAT_TOL * np.maximum(1.0, np.abs(want[y:y + 256].astype(np.float64)))
        bad[y:y + 256] = (d > lim).any(axis=-1)
        worst = max(worst, float(d.max()))
    if keep:
        ys, xs = np.nonzero(bad)
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/fullsize_mismatch_%d%s.npz" % (size, "_shadow" if shadow else ""), ys=ys, xs=xs, got=got[ys, xs], want=want[ys, xs])
    return int(bad.sum()), worst


n, worst = mismatch(script, keep=True)
print("size", size, "shadow", shadow, "pixels off", n, "max |diff|", worst)
ops = split(script)
draw_ends = [e for (name, b, e) in ops if name in DRAWS]
lo, hi = 0, len(draw_ends) - 1          # smallest prefix (in draws) that mismatches
if n:
    while lo < hi:
        mid = (lo + hi) // 2
        k, _ = mismatch(script[:draw_ends[mid]])
        print("  draws 0..%d -> %d pixels off" % (mid, k), flush=True)
        if k: hi = mid
        else: lo = mid + 1
    print("first draw that disagrees: #%d" % lo)
    prev = draw_ends[lo - 1] if lo else 0
    for (nm, b, e) in ops:
        if prev <= b < draw_ends[lo]:
            print("    ", nm, show(script, nm, b, e))
    # that draw alone (with the state in force) on an empty canvas
    state = b"".join(script[b:e] for (nm, b, e) in ops if e <= prev and (nm.startswith("SET_") or nm in ("TRANSLATE", "ROTATE", "SCALE")))
    k, w = mismatch(state + script[prev:draw_ends[lo]])
    print("that draw alone on an empty canvas: %d pixels off, max %.3g" % (k, w))
