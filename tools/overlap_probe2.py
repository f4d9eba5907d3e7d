#!/usr/bin/env python
"""Is the concurrent-replay throughput host-bound?  T host threads x L canvases each."""
import ctypes as C, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import harness as H
from canvas_ity_b200 import _native
lib = _native.load()
size = 4096
frame = H.lower_script(H.tiger_script(size, size), size, size)[0]
def make(n):
    out = []
    for _ in range(n):
        cv = C.c_void_p()
        assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
        assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
        lib.cb200_set_stage_timing(cv, 0)
        out.append(cv)
    return out
for threads, per in ((1, 8), (2, 4), (4, 2), (8, 1)):
    groups = [make(per) for _ in range(threads)]
    def run(cvs, rounds):
        for _ in range(rounds):
            for cv in cvs: lib.cb200_frame_replay(cv, 1)
        for cv in cvs: lib.cb200_sync(cv)
    for g in groups: run(g, 3)
    rounds = 16
    ts = [threading.Thread(target=run, args=(g, rounds)) for g in groups]
    t0 = time.perf_counter()
    [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.perf_counter() - t0
    n = rounds * threads * per
    print("threads %d x canvases %d: %.1f frames/s" % (threads, per, n / dt), flush=True)
    # host cost of queueing alone
    t0 = time.perf_counter()
    for cv in groups[0]: lib.cb200_frame_replay(cv, 1)
    q = (time.perf_counter() - t0) / per
    for cv in groups[0]: lib.cb200_sync(cv)
    print("   host time to queue one frame: %.0f us" % (q * 1e6), flush=True)
    for g in groups:
        for cv in g: lib.cb200_canvas_destroy(cv)
