#!/usr/bin/env python
"""GPU: bit-level comparison of the device's stroke outline with the oracle's on small probe scenes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H

lib, orc = H.product_library(), H.oracle_library()
W = 200


def probe(name, script):
    fr = H.lower_script(script, W, W)[0]
    want = np.zeros((200000, 4), np.float32)
    n = orc.oracle_debug_edges(C.addressof(fr.frame), fr.n_draws - 1, want.ctypes.data, len(want))
    want = want[:n]
    want = want[np.abs(want[:, 3] - want[:, 1]) >= 2.0e-5]
    cv = C.c_void_p()
    assert lib.cb200_canvas_create(W, W, 0, C.byref(cv)) == 0
    assert lib.cb200_submit(cv, C.byref(fr.frame)) == 0, lib.cb200_last_error()
    got = np.zeros((200000, 4), np.float32)
    m = lib.cb200_debug_lines(cv, got.ctypes.data, None, len(got))
    got = got[:m]
    lib.cb200_canvas_destroy(cv)
    a = set(map(bytes, want.view(np.uint8).reshape(len(want), 16)))
    b = set(map(bytes, got.view(np.uint8).reshape(len(got), 16)))
    print("%s: oracle %d edges, device %d; only oracle %d, only device %d" % (name, len(want), len(got), len(a - b), len(b - a)))
    oa = np.array([np.frombuffer(e, np.float32) for e in sorted(a - b)][:6])
    ob = np.array([np.frombuffer(e, np.float32) for e in sorted(b - a)][:6])
    if len(oa): print("   oracle:\n", oa, "\n   device:\n", ob)
    np.savez("gpurun_out/join_%s.npz" % name, want=want, got=got)


def stroke(join, cap, lw, pts, closed=False):
    w = H.ScriptWriter()
    w.floats("SET_LINE_WIDTH", lw); w.ints("SET_LINE_JOIN", join); w.ints("SET_LINE_CAP", cap)
    w.bare("BEGIN_PATH"); w.floats("MOVE_TO", *pts[0])
    for p in pts[1:]: w.floats("LINE_TO", *p)
    if closed: w.bare("CLOSE_PATH")
    w.bare("STROKE")
    return w.take()


os.makedirs("gpurun_out", exist_ok=True)
zig = [(20.3, 30.1), (90.7, 45.2), (60.2, 120.9), (150.4, 100.3), (120.8, 170.6)]
probe("miter", stroke(0, 0, 9.0, zig))
probe("bevel", stroke(1, 0, 9.0, zig))
probe("round", stroke(2, 0, 9.0, zig))
probe("round_thin", stroke(2, 0, 1.0, zig))
probe("round_closed", stroke(2, 0, 9.0, zig, closed=True))
probe("round_cap", stroke(0, 2, 9.0, zig))
w = H.ScriptWriter()
w.floats("SET_LINE_WIDTH", 1.0); w.ints("SET_LINE_JOIN", 2)
w.floats("SET_FONT", 40.0); w.raw("B", 1); w.blob(H.font_a())
w.floats("STROKE_TEXT", 10.0, 100.0, 1.0e30); w.blob(b"CDE")
probe("text_round", w.take())
w = H.ScriptWriter()
w.floats("SET_LINE_WIDTH", 1.0); w.ints("SET_LINE_JOIN", 0)
w.floats("SET_FONT", 40.0); w.raw("B", 1); w.blob(H.font_a())
w.floats("STROKE_TEXT", 10.0, 100.0, 1.0e30); w.blob(b"CDE")
probe("text_miter", w.take())
