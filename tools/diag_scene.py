#!/usr/bin/env python
"""Delta-debugging helper for a failing fuzz seed (GPU): find the first draw after which the CUDA back end and
the oracle disagree, print that draw's ops, then try dropping single ops of the scene to see which matter."""
import os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tests.test_random_scenes import random_scene, random_scene_wide, SIZE
from canvas_ity_b200.script import OPS

from tools.diag_scene_lib import split, show, DRAWS, FMT, INTS

wide = '--wide' in sys.argv
seed = int([a for a in sys.argv[1:] if not a.startswith('--')][0])
lib = H.product_library()
if wide: s, W, Hh = random_scene_wide(seed)
else: s, W, Hh = random_scene(seed), SIZE, SIZE
ops = split(s)


def nbad(sc):
    got, want = H.render_script(lib, sc, W, Hh), H.render_oracle(sc, W, Hh)
    return H.float_mismatch(got['f32'], want['f32'])[0]


print('seed', seed, 'canvas', W, Hh, 'full nbad', nbad(s))
prev = 0
for k, (name, b, e) in enumerate(ops):
    if name not in DRAWS: continue
    n = nbad(s[:e])
    if n:
        print('first failing draw ends at op', k, name, 'nbad', n)
        for (nm, bb, ee) in ops:
            if prev <= bb < e: print('    ', nm, show(s, nm, bb, ee))
        break
    prev = e
print('state ops before that draw that are still in force (last of each kind):')
last = {}
for (nm, bb, ee) in ops:
    if bb >= prev: break
    if nm.startswith('SET_') or nm in ('TRANSLATE', 'ROTATE', 'SCALE', 'SAVE', 'RESTORE', 'CLIP'): last[nm] = show(s, nm, bb, ee)
print('    ', last)
print('dropping single state ops (whole scene):')
for k, (nm, bb, ee) in enumerate(ops):
    if not (nm.startswith('SET_') or nm in ('TRANSLATE', 'ROTATE', 'SCALE', 'CLIP')): continue
    if bb >= e: break
    r = nbad(s[:bb] + s[ee:e])
    print('   drop %3d %-22s %s -> nbad %d' % (k, nm, show(s, nm, bb, ee), r))
