#!/usr/bin/env python
"""Seed sweep over tests/test_random_scenes.random_scene: --mode gpu compares the CUDA back end with the
oracle (needs a GPU), --mode ref compares the oracle with the reference build (CPU).  Prints failing seeds."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tests.test_random_scenes import random_scene, random_scene_wide, integer_scene, SIZE

ap = argparse.ArgumentParser()
ap.add_argument("--mode", choices=["gpu", "ref"], default="gpu")
ap.add_argument("--first", type=int, default=24)
ap.add_argument("--count", type=int, default=200)
ap.add_argument("--wide", action="store_true", help="the second generator (odd sizes, far-away geometry, images, text layout)")
ap.add_argument("--integer", action="store_true", help="the third generator (shadowed polygons on integer coordinates)")
a = ap.parse_args()
lib = H.product_library() if a.mode == "gpu" else H.reference_library()
bad = []
for seed in range(a.first, a.first + a.count):
    script, w, h = random_scene_wide(seed) if a.wide else integer_scene(seed) if a.integer else (random_scene(seed), SIZE, SIZE)
    try:
        got = H.render_script(lib, script, w, h) if a.mode == "gpu" else H.render_oracle(script, w, h)
        want = H.render_oracle(script, w, h) if a.mode == "gpu" else H.render_script(lib, script, w, h)
        nbad, worst = H.float_mismatch(got["f32"], want["f32"])
        n8 = H.rgba8_mismatch(got["rgba8"], want["rgba8"])[2]
    except Exception as e:                      # noqa: BLE001
        nbad, worst, n8 = -1, 0.0, -1
        print("seed", seed, "ERROR", e)
    if nbad or n8:
        bad.append(seed)
        print("seed %d: %d floats off (max %.3g), %d pixels beyond 1 LSB" % (seed, nbad, worst, n8), flush=True)
print("%s sweep %d..%d: %d failing seeds %s" % (a.mode, a.first, a.first + a.count - 1, len(bad), bad))
