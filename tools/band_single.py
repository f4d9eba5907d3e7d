#!/usr/bin/env python
"""One band of the 4096^2 tiger on one GPU (for ncu launch lists): BAND=y0,rows"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import harness as H
from canvas_ity_b200 import _native
lib = _native.load()
size = 4096
y0, rows = [int(v) for v in os.environ.get("BAND", "0,2048").split(",")]
frame = H.lower_script(H.tiger_script(size, size), size, size)[0]
cv = C.c_void_p()
assert lib.cb200_canvas_create_band(size, size, y0, rows, 0, C.byref(cv)) == 0
assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
st = _native.Stats()
for i in range(4):
    assert lib.cb200_frame_replay(cv, 1) == 0
    lib.cb200_get_stats(cv, C.byref(st))
print("band [%d,%d): frame %.3f geometry %.3f raster %.3f sort %.3f coverage %.3f composite %.3f runs %d tiles %d" %
      (y0, y0 + rows, st.last_frame_ms, st.geometry_ms, st.raster_ms, st.sort_ms, st.coverage_ms, st.composite_ms, st.raw_runs, st.tile_entries))
