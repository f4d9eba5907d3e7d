#!/usr/bin/env python
"""Config 4 of BASELINE.json: full-canvas fills at 8192x8192 with solid / linear / radial /
bicubic-image brushes under copy, xor, lighter and source_over; reports the compositor's CUDA-event
time and its algorithmic bandwidth (32 B per composited pixel) against the measured HBM peak."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from canvas_ity_b200 import _native
from canvas_ity_b200.script import ScriptWriter

SIZE = int(os.environ.get("FILL_SIZE", 8192))
OPS = {"source_copy": 2, "exclusive_or": 15, "lighter": 10, "source_over": 14}


def main():
    lib = _native.load()
    peak = json.load(open(os.path.join(H.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(H.ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    kinds = os.environ.get("FILL_KINDS", "solid,linear,radial,image").split(",")
    ops = {k: v for k, v in OPS.items() if k in os.environ.get("FILL_OPS", ",".join(OPS)).split(",")}
    for kind in kinds:
        for opname, op in ops.items():
            # background first so that the measured draw blends over real pixels
            bg = ScriptWriter(); bg.ints("SET_COLOR", 0); bg.raw("4f", 0.9, 0.8, 0.1, 0.6); bg.floats("FILL_RECTANGLE", 0, 0, float(SIZE), float(SIZE))
            frame = H.lower_script(H.config4_script(kind, op, SIZE), SIZE, SIZE)[0]
            bgf = H.lower_script(bg.take(), SIZE, SIZE)[0]
            cv = C.c_void_p()
            assert lib.cb200_canvas_create(SIZE, SIZE, 0, C.byref(cv)) == 0
            assert lib.cb200_submit(cv, C.byref(bgf.frame)) == 0
            assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
            st = _native.Stats()
            ts = []
            for i in range(8):
                assert lib.cb200_frame_replay(cv, 0) == 0
                assert lib.cb200_get_stats(cv, C.byref(st)) == 0
                if i >= 3:
                    ts.append((st.composite_ms, st.last_frame_ms))
            comp = float(np.mean([t[0] for t in ts])); frame_ms = float(np.mean([t[1] for t in ts]))
            px = int(st.composited_pixels)
            gbs = px * 32 / (comp * 1e-3) / 1e9
            rows.append(dict(brush=kind, op=opname, composite_ms=comp, frame_ms=frame_ms, composited_px=px,
                             algorithmic_GBs=gbs, frac_of_measured_hbm=gbs / peak))
            print("%-7s %-13s composite %.3f ms  frame %.3f ms  px %d  %.0f GB/s  %.2f of HBM peak" %
                  (kind, opname, comp, frame_ms, px, gbs, gbs / peak), flush=True)
            lib.cb200_canvas_destroy(cv)
    json.dump(rows, open(os.path.join(H.ROOT, "gpurun_out", "fill_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
