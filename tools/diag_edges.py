#!/usr/bin/env python
"""GPU: compare the outline edges the device produces for single draws of the tiger (flatten / stroke kernels,
cb200_debug_lines) with the oracle's (oracle_debug_edges), bit for bit.  usage: diag_edges.py size [draw ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tools.diag_scene_lib import split, DRAWS

size = int(sys.argv[1])
which = [int(a) for a in sys.argv[2:]]
script = H.tiger_script(size, size)
ops = split(script)
draw_ends = [e for (name, b, e) in ops if name in DRAWS]
lib, orc = H.product_library(), H.oracle_library()
for di in (which or range(len(draw_ends))):
    prev = draw_ends[di - 1] if di else 0
    state = b"".join(script[b:e] for (nm, b, e) in ops if e <= prev and (nm.startswith("SET_") or nm in ("TRANSLATE", "ROTATE", "SCALE")))
    one = state + script[prev:draw_ends[di]]
    fr = H.lower_script(one, size, size)[0]
    want = np.zeros((400000, 4), np.float32)
    n = orc.oracle_debug_edges(C.addressof(fr.frame), 0, want.ctypes.data, len(want))
    want = want[:n]
    want = want[np.abs(want[:, 3] - want[:, 1]) >= 2.0e-5]
    cv = C.c_void_p()
    assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
    assert lib.cb200_submit(cv, C.byref(fr.frame)) == 0, lib.cb200_last_error()
    got = np.zeros((400000, 4), np.float32)
    m = lib.cb200_debug_lines(cv, got.ctypes.data, None, len(got))
    got = got[:m]
    lib.cb200_canvas_destroy(cv)
    a = set(map(bytes, want.view(np.uint8).reshape(len(want), 16)))
    b = set(map(bytes, got.view(np.uint8).reshape(len(got), 16)))
    if a != b:
        print("draw %d: oracle %d edges, device %d; only oracle %d, only device %d" % (di, len(want), len(got), len(a - b), len(b - a)))
        for e in sorted(a - b)[:8]: print("   oracle:", np.frombuffer(e, np.float32))
        for e in sorted(b - a)[:8]: print("   device:", np.frombuffer(e, np.float32))
    elif which:
        print("draw %d: %d edges identical" % (di, len(want)))
print("done")
