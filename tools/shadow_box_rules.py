#!/usr/bin/env python
"""CPU-only: how often does a per-edge rule for the shadow working rectangle reproduce the box the reference
derives from its polygon clip (oracle_debug_shadow_boxes)?  Sweeps the fuzz scenes' shadowed draws."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import harness as H
from tests.test_random_scenes import random_scene, random_scene_wide, SIZE

orc = H.oracle_library()
orc.oracle_debug_shadow_boxes.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 8)]
orc.oracle_debug_shadow_boxes.restype = C.c_int
DRAW_BYTES = 140
import struct
NA = int(sys.argv[1]) if len(sys.argv) > 1 else 400
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 900
FIRST = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for rule in (2, 0, 1):          # 2 = pairs + clip order (left, top, right, bottom); 0 = pairs; 1 = today's CUDA rule (modelled)
    bad, total, bad_seeds = 0, 0, []
    for gen, count in (("a", NA), ("b", NB)):
        for seed in range(FIRST, FIRST + count):
            if gen == "a": script, w, h = random_scene(seed), SIZE, SIZE
            else: script, w, h = random_scene_wide(seed)
            for fr in H.lower_script(script, w, h):
                draws = bytes(fr.parts["draws"])
                for di in range(fr.n_draws):
                    rec = draws[di * DRAW_BYTES:(di + 1) * DRAW_BYTES]
                    kind = struct.unpack_from("<I", rec, 0)[0]
                    shadow = struct.unpack_from("<7f", rec, 108)            # shadow_color[4], offset x, y, blur
                    if kind == 2 or shadow[3] == 0.0 or (shadow[4] == 0.0 and shadow[5] == 0.0 and shadow[6] == 0.0):
                        continue
                    out = (C.c_int * 8)()
                    orc.oracle_debug_shadow_boxes(C.addressof(fr.frame), di, w, h, rule, C.byref(out))
                    total += 1
                    if list(out[0:4]) != list(out[4:8]):
                        bad += 1
                        if (gen, seed) not in bad_seeds: bad_seeds.append((gen, seed))
    print("rule %d: %d of %d shadowed draws differ; scenes %s" % (rule, bad, total, bad_seeds[:40]))
