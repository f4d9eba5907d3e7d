#!/usr/bin/env python
"""Delta-debugging helper for a failing fuzz seed (GPU): find the first draw after which the CUDA back end and
the oracle disagree, print that draw's ops, then try dropping single ops of the scene to see which matter."""
import os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tests.test_random_scenes import random_scene, random_scene_wide, SIZE
from canvas_ity_b200.script import OPS

FMT = {'SCALE': 8, 'ROTATE': 4, 'TRANSLATE': 8, 'SET_GLOBAL_ALPHA': 4, 'SET_COMPOSITE': 4, 'SET_SHADOW_COLOR': 16, 'SET_SHADOW_OFFSET_X': 4,
       'SET_SHADOW_OFFSET_Y': 4, 'SET_SHADOW_BLUR': 4, 'SET_LINE_WIDTH': 4, 'SET_LINE_CAP': 4, 'SET_LINE_JOIN': 4, 'SET_MITER_LIMIT': 4,
       'SET_LINE_DASH_OFFSET': 4, 'SET_COLOR': 20, 'SET_LINEAR_GRADIENT': 20, 'SET_RADIAL_GRADIENT': 28, 'ADD_COLOR_STOP': 24, 'BEGIN_PATH': 0,
       'MOVE_TO': 8, 'CLOSE_PATH': 0, 'LINE_TO': 8, 'QUADRATIC_CURVE_TO': 16, 'BEZIER_CURVE_TO': 24, 'ARC': 24, 'ARC_TO': 20, 'RECTANGLE': 16,
       'FILL': 0, 'STROKE': 0, 'CLIP': 0, 'FILL_RECTANGLE': 16, 'STROKE_RECTANGLE': 16, 'CLEAR_RECTANGLE': 16, 'SAVE': 0, 'RESTORE': 0,
       'SET_TEXT_ALIGN': 4, 'SET_TEXT_BASELINE': 4}
INTS = ('SET_COMPOSITE', 'SET_LINE_CAP', 'SET_LINE_JOIN', 'SET_TEXT_ALIGN', 'SET_TEXT_BASELINE')
DRAWS = ('FILL', 'STROKE', 'FILL_RECTANGLE', 'STROKE_RECTANGLE', 'CLEAR_RECTANGLE', 'FILL_TEXT', 'STROKE_TEXT', 'CLIP', 'DRAW_IMAGE', 'PUT_IMAGE_DATA')


def split(s):
    ops, i = [], 0
    while i < len(s):
        name, j = OPS[s[i]], i + 1
        if name in FMT: j += FMT[name]
        elif name == 'SET_LINE_DASH': j += 4 + 4 * struct.unpack_from('<i', s, j)[0]
        elif name == 'SET_PATTERN': j += 20; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'SET_FONT': j += 5; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name in ('FILL_TEXT', 'STROKE_TEXT'): j += 12; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'DRAW_IMAGE': j += 28; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'PUT_IMAGE_DATA': j += 20; j += 4 + struct.unpack_from('<I', s, j)[0]
        else: raise SystemExit('unknown op ' + name)
        ops.append((name, i, j)); i = j
    return ops


def show(s, name, b, e):
    if name in INTS: return struct.unpack_from('<i', s, b + 1)
    if name in FMT and FMT[name]:
        v = struct.unpack_from('<%df' % (FMT[name] // 4), s, b + 1)
        return tuple(round(x, 2) for x in v)
    if name == 'SET_LINE_DASH': return struct.unpack_from('<i', s, b + 1)
    return ''


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
