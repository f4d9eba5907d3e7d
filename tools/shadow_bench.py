#!/usr/bin/env python
"""Config 3 of BASELINE.json: tiger at 4096x4096 with global_alpha 0.9, shadow_blur 16, shadow alpha 0.5."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from canvas_ity_b200 import _native

size = int(os.environ.get("SIZE", 4096))
lib = _native.load()
script = H.tiger_script(size, size, global_alpha=0.9, shadow_blur=16.0, shadow_color=(0, 0, 0, 0.5))
frame = H.lower_script(script, size, size)[0]
cv = C.c_void_p()
assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0, lib.cb200_last_error()
st = _native.Stats()
ts = []
for i in range(8):
    assert lib.cb200_frame_replay(cv, 1) == 0, lib.cb200_last_error()
    assert lib.cb200_get_stats(cv, C.byref(st)) == 0, lib.cb200_last_error()
    ts.append(st.last_frame_ms)
print("tiger %d shadow: frame ms %s  fps %.1f  composited %d  plane floats %d  tiles %d runs %d" %
      (size, ["%.2f" % t for t in ts], 1e3 / min(ts[3:]), st.composited_pixels, st.shadow_pixels, st.tile_entries, st.raw_runs))
print("stages: geometry %.3f raster %.3f sort %.3f composite %.3f" % (st.geometry_ms, st.raster_ms, st.sort_ms, st.composite_ms))
lib.cb200_canvas_destroy(cv)
