#!/usr/bin/env python
"""Throughput of L canvases replaying the tiger frame concurrently (one stream each, one host thread)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import harness as H
from canvas_ity_b200 import _native
lib = _native.load()
size = 4096
frame = H.lower_script(H.tiger_script(size, size), size, size)[0]
for lanes in (1, 2, 4, 8):
    cvs = []
    for _ in range(lanes):
        cv = C.c_void_p()
        assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
        assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
        lib.cb200_set_stage_timing(cv, 0)
        cvs.append(cv)
    for _ in range(5):
        for cv in cvs: assert lib.cb200_frame_replay(cv, 1) == 0
    for cv in cvs: lib.cb200_sync(cv)
    rounds = 60 // lanes
    t0 = time.perf_counter()
    for _ in range(rounds):
        for cv in cvs: lib.cb200_frame_replay(cv, 1)
    for cv in cvs: lib.cb200_sync(cv)
    dt = time.perf_counter() - t0
    print("lanes %d: %.1f frames/s (%.3f ms per frame)" % (lanes, rounds * lanes / dt, dt / (rounds * lanes) * 1e3), flush=True)
    for cv in cvs: lib.cb200_canvas_destroy(cv)
