#!/usr/bin/env python
"""Config 5 of BASELINE.json: a batch of independent 256x256 canvases (random paths + glyph text)
rendered as one device frame.  Reports canvases/s from the CUDA-event frame time and the stage split."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from canvas_ity_b200 import _native

lib = _native.load()
for n in [int(a) for a in (sys.argv[1:] or ["1024"])]:
    t0 = time.time()
    scripts = [H.config5_script(i) for i in range(n)]
    t1 = time.time()
    batch = lib.cv_batch_create(n, 256, 256, 0)
    assert batch, lib.cv_last_error()
    for rep in range(3):
        cv = lib.cv_batch_device(batch)
        assert lib.cb200_clear(cv) == 0
        t2 = time.time()
        for i, s in enumerate(scripts):
            lib.cv_run_script(lib.cv_batch_canvas(batch, i), s, len(s), None, 0, None)
        t3 = time.time()
        assert lib.cv_batch_flush(batch) == 0, lib.cv_last_error()
        t4 = time.time()
        st = _native.Stats()
        assert lib.cb200_get_stats(cv, C.byref(st)) == 0
        print("n=%d rep %d: device frame %.2f ms -> %.0f canvases/s | geometry %.2f raster %.2f sort %.2f composite %.2f | "
              "host: scripts %.2fs lower %.2fs flush+submit %.2fs | draws %d runs %d tiles %d composited %d launches %d" %
              (n, rep, st.last_frame_ms, n / (st.last_frame_ms * 1e-3), st.geometry_ms, st.raster_ms, st.sort_ms, st.composite_ms,
               t1 - t0, t3 - t2, t4 - t3, st.draws, st.raw_runs, st.tile_entries, st.composited_pixels, st.kernel_launches), flush=True)
    # spot check against solo renders
    for i in (0, n // 2, n - 1):
        got = np.zeros((256, 256, 4), np.float32)
        assert lib.cv_batch_read_f32(batch, i, got.ctypes.data) == 0
        alone = H.render_script(lib, scripts[i], 256, 256)["f32"]
        assert np.array_equal(got.view(np.uint32), alone.view(np.uint32)), i
    print("spot checks vs solo renders: identical")
    lib.cv_batch_destroy(batch)
