#!/usr/bin/env python
"""fuzz_scenes.py for an explicit list of seeds of the second generator: python tools/fuzz_seeds.py 106 145 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import harness as H
from tests.test_random_scenes import random_scene_wide
lib = H.product_library()
for seed in [int(a) for a in sys.argv[1:]]:
    script, w, h = random_scene_wide(seed)
    got, want = H.render_script(lib, script, w, h), H.render_oracle(script, w, h)
    nbad, worst = H.float_mismatch(got["f32"], want["f32"])
    print("seed %d: %d floats off (max %.3g), %d pixels beyond 1 LSB" % (seed, nbad, worst, H.rgba8_mismatch(got["rgba8"], want["rgba8"])[2]))
