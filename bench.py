#!/usr/bin/env python
"""bench.py -- headline benchmark: Ghostscript tiger (305 draws, 2380 cubics) at 4096 x 4096 (BASELINE.json `tiger_4096`).

Contract (see the round brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.

  step     one whole frame on EACH of `--lanes` canvases (config.frames_per_step): fresh canvas + flatten/stroke + scan
           conversion + sort + coverage + tile compositing of all 305 draws (the reference demo's timed region,
           tiger.cpp:104-4323: draw calls only, fresh canvas each frame; its readback is outside the timed region).
  value    frames/s with the lowered frame already resident in HBM (cb200_frame_upload once, cb200_frame_replay(clear=1)
           per frame, queued back to back) on `--lanes` canvases that replay concurrently (one CUDA stream each).  A timed
           region is EXACTLY K steps bracketed by a barrier + synchronize; its span is taken with CUDA events on the canvas
           streams, from the earliest begin event to the latest end event.  Every stream has one untimed frame queued
           before its begin event is recorded, so the host's enqueue latency is not inside the span.  The region is
           repeated until at least 0.6 s have been timed; `value` comes from the MEDIAN span (max over ranks), the
           spread is reported under `span_ms`.
  e2e      frames/s through the public drop-in API with HOST buffers every frame: canvas-script replay (host path
           building + lowering) -> pinned H2D -> kernels -> get_image_data (sRGB/dither kernel + D2H of the RGBA8 image).
  roofline the tile compositor (k_composite): algorithmic bytes = 32 B per composited pixel / its CUDA-event
           time, against the measured HBM peak; `dram_frac` = the kernel's real DRAM traffic (committed ncu capture) /
           time / peak, which is what the kernel actually asks of HBM.
  passes   the other passes and configurations BASELINE.json names, each against its own algorithmic bytes: readback,
           sort, coverage, PNG encode, bulk hit test, config 3 (shadows), config 4 (8192^2 full-canvas fills), config 5
           (a batch of 256^2 canvases); with N > 1 also config 3 as ONE frame sharded by scanline bands + NCCL all_gather
           (strong scaling) and config 5 split over the ranks.
  cpu_baseline  the unmodified reference (oracle/_ref/libcanvas_ref_fast.so, -O3 -march=native -ffp-contract=off)
           rendering the same call stream on one host thread, bounded sample.

N > 1 (torchrun): frames are independent canvases, every rank renders whole frames (weak scaling, no data-path
collective).  `--impl reference` times the reference's own CPU implementation with all host threads: one step = one frame
per host thread.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tiger_4096_frames_per_s"
UNIT = "frames/s"
SIZE = 4096
ALGO_BYTES_PER_COMPOSITED_PIXEL = 32.0     # 16 B load + 16 B store of the float4 texel
MIN_TIMED_S = 0.6


def workload_name(size):
    """config.workload -- the same string in both arms."""
    return "tiger_%d: demos/tiger call stream (305 draws) fit to %dx%d, source_over, no shadow, fresh canvas per frame" % (size, size, size)


def ncu_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture
    (profiles/*ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None if absent."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_traffic.json")), reverse=True):
        entry = json.load(open(path)).get(kernel)
        if entry:
            return entry["dram_bytes_per_launch"]
    return None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling recipe's line)."""

    def __init__(self, index, enabled=True):
        self.enabled = enabled
        self.rows, self.stop = [], threading.Event()
        self.cmd = ["nvidia-smi", "-i", str(index),
                    "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                    "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                    "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"]
        self.proc = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ reference arm ----

def cpu_renderer(size):
    """(kind, frame) -- `frame()` renders one whole tiger frame on the calling CPU thread: the
    unmodified reference from oracle/_ref when it was built (kind "reference"), else the oracle port
    fed with the frame lowered once by the front end (kind "port").  Draw calls only, fresh canvas per
    frame: the timed region of demos/tiger/tiger.cpp:104-4323."""
    from tests import harness as H
    script = H.tiger_script(size, size)
    lib = H.reference_library(fast=True) or H.reference_library()
    if lib is not None:
        def frame():
            h = lib.cv_create(size, size)
            lib.cv_run_script(h, script, len(script), None, 0, None)
            lib.cv_destroy(h)
        return "reference", frame
    orc = H.oracle_library()
    lowered = H.lower_script(script, size, size)

    def frame():
        o = orc.oracle_canvas_create(size, size)
        for f in lowered:
            orc.oracle_submit(o, C.byref(f.frame))
        orc.oracle_canvas_destroy(o)
    return "port", frame


def run_reference(args, rank, world):
    """The reference's own CPU renderer on the same call stream, all host threads.  One step = one frame per host
    thread (ctypes releases the GIL), W warm-up steps, exactly K timed steps."""
    if rank != 0:
        return
    size = args.size
    kind, frame = cpu_renderer(size)
    threads = os.cpu_count() or 1

    def step():
        ts = [threading.Thread(target=frame) for _ in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    spans = [step() for _ in range(args.steps)]
    t = sum(spans)
    fps = args.steps * threads / t
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(size), "canvas": [size, size], "frames_per_step": threads,
                       "l2": "n/a (CPU arm)"},
            "step_ms": {"median": 1e3 * float(np.median(spans)), "min": 1e3 * min(spans), "max": 1e3 * max(spans)},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d steps x %d concurrent frames (one per host thread), draw calls only" % (args.steps, threads)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline_sample(size, budget_s=15.0):
    """Single-thread CPU frames for about `budget_s` seconds (what one frame costs the reference)."""
    kind, frame = cpu_renderer(size)
    n, t0 = 0, time.perf_counter()
    best = 1e30
    while True:
        t1 = time.perf_counter()
        frame()
        best = min(best, time.perf_counter() - t1)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 24:
            break
    return {"value": 1.0 / best, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d whole %dx%d tiger frames on one host thread (the reference is single-threaded), best-of" % (n, size, size)}


# ------------------------------------------------------------------------ our arm ----

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=8, help="canvases replaying concurrently per GPU in the device-resident arm (= frames per step)")
    ap.add_argument("--e2e-lanes", type=int, default=0,
                    help="canvases kept in flight per rank by the end-to-end arm (default: up to 4, one host thread each, "
                         "as the host cores allow)")
    ap.add_argument("--skip-configs", action="store_true", help="leave out the config 3 / 4 / 5 passes (contract test on small canvases)")
    ap.add_argument("--fill-size", type=int, default=8192, help="canvas of the config 4 fills")
    ap.add_argument("--batch", type=int, default=2048, help="canvases per GPU in the config 5 pass")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if args.e2e_lanes <= 0:
        args.e2e_lanes = max(2, min(4, (os.cpu_count() or 4) // max(1, world)))

    import torch
    import torch.distributed as dist
    from tests import harness as H
    from canvas_ity_b200 import _native, sharding
    lib = _native.load()
    if lib.cb200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; this back end has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    pin_to_gpu_numa_node(local, world)
    size = args.size
    peak_gbs, peak_src = measured_hbm_peak()

    def check(rc):
        if rc:
            raise RuntimeError(lib.cb200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def new_canvas(w, h, y0=0, rows=None):
        cv = C.c_void_p()
        check(lib.cb200_canvas_create_band(w, h, y0, h if rows is None else rows, local, C.byref(cv)))
        return cv

    stats = _native.Stats()
    script = H.tiger_script(size, size)
    frame = H.lower_script(script, size, size)[0]

    def timed_regions(cvs, steps, min_seconds, max_regions=64, after_frame=None, all_ranks=True):
        """Spans (ms) of repeated regions of exactly `steps` steps (one frame on every canvas per step).  `all_ranks`:
        every rank is in here (barriers between regions); False for passes that only rank 0 runs."""
        spans = []
        ms = C.c_float()
        sync = barrier if all_ranks else torch.cuda.synchronize
        n_regions = 3                                       # fixed after the first region (the same on every rank)
        while len(spans) < n_regions:
            for c in cvs:                                   # one untimed frame: the streams are busy when the
                check(lib.cb200_frame_replay(c, 1))         # begin events are recorded
            if all_ranks and world > 1 and not spans:
                dist.barrier()
            for c in cvs:
                check(lib.cb200_timer_begin(c))
            for _ in range(steps):
                for c in cvs:
                    check(lib.cb200_frame_replay(c, 1))     # fresh canvas + all kernels of the frame, queued async
                    if after_frame:
                        after_frame(c)
            for c in cvs:
                check(lib.cb200_timer_stop(c))
            span = 0.0
            for c in cvs:
                check(lib.cb200_timer_end(c, C.byref(ms), None, None))
            for a in cvs:                                   # earliest begin to latest end
                for b in cvs:
                    check(lib.cb200_timer_between(a, b, C.byref(ms)))
                    span = max(span, ms.value)
            sync()
            spans.append(span)
            if len(spans) == 1:
                first = torch.tensor([span], dtype=torch.float64, device="cuda")
                if all_ranks and world > 1:
                    dist.all_reduce(first, op=dist.ReduceOp.MAX)
                n_regions = int(min(max_regions, max(3, np.ceil(min_seconds * 1e3 / max(float(first[0]), 1e-3)))))
        return spans

    # ---- device-resident arm: `value` ----
    # `lanes` canvases (one CUDA stream and one set of work buffers each) replay the resident frame concurrently: the
    # latency-bound geometry / sort / coverage kernels of one frame overlap the compositor of another.
    n_canvases = max(1, args.lanes)
    cvs = [new_canvas(size, size) for _ in range(n_canvases)]
    cv = cvs[0]
    for c in cvs:
        check(lib.cb200_frame_upload(c, C.byref(frame.frame)))
        check(lib.cb200_set_stage_timing(c, 0))             # only the frame and the compositor carry CUDA events
    for _ in range(args.warmup):
        for c in cvs:
            check(lib.cb200_frame_replay(c, 1))
    launches_before = 0
    for c in cvs:
        check(lib.cb200_sync(c))
        check(lib.cb200_get_stats(c, C.byref(stats)))
        launches_before += stats.kernel_launches
    barrier()
    with ClockSampler(local, enabled=rank == 0) as clocks:      # one nvidia-smi poller per job, not per rank
        spans = timed_regions(cvs, args.steps, MIN_TIMED_S)
    launches = -launches_before
    graph_replays = 0
    for c in cvs:
        check(lib.cb200_get_stats(c, C.byref(stats)))
        launches += int(stats.kernel_launches)
        graph_replays += int(stats.graph_replays)
    launches_per_region = launches // len(spans)
    composited = int(stats.composited_pixels)
    span_ms = float(np.median(spans))

    # ---- the compositor alone on the GPU: one canvas, stream launches (every frame then carries its own compositor
    # events), enough frames for a stable mean ----
    for extra in cvs[1:]:
        lib.cb200_canvas_destroy(extra)
    check(lib.cb200_set_graph_replay(cv, 0))
    barrier()
    elapsed_ms, comp_sum_ms, comp_frames = C.c_float(), C.c_float(), C.c_uint32()
    single_frames = max(args.steps, 64)
    check(lib.cb200_frame_replay(cv, 1))
    check(lib.cb200_timer_begin(cv))
    for _ in range(single_frames):
        check(lib.cb200_frame_replay(cv, 1))
    check(lib.cb200_timer_end(cv, C.byref(elapsed_ms), C.byref(comp_sum_ms), C.byref(comp_frames)))
    single_s = elapsed_ms.value / 1e3
    comp_ms = comp_sum_ms.value / max(1, comp_frames.value)
    check(lib.cb200_set_stage_timing(cv, 1))
    split = []
    for _ in range(7):
        check(lib.cb200_frame_replay(cv, 1))
        check(lib.cb200_get_stats(cv, C.byref(stats)))
        split.append((stats.geometry_ms, stats.raster_ms, stats.sort_ms, stats.coverage_ms, stats.composite_ms, stats.last_frame_ms))
    stages_ms = dict(zip(("geometry", "raster", "sort", "coverage", "composite", "frame_with_stage_events"),
                         [float(np.median(c)) for c in zip(*split[2:])]))
    runs = int(stats.raw_runs)

    # ---- the other passes of the north star, each against its own algorithmic bytes ----
    passes = {}
    if rank == 0:
        rb = []
        dev_ptr = C.c_void_p()
        for _ in range(8):                                      # sRGB + dither kernel alone, device to device
            check(lib.cb200_read_rgba8_device(cv, C.byref(dev_ptr)))
            check(lib.cb200_get_stats(cv, C.byref(stats)))
            rb.append(stats.readback_ms)
        rb_ms = float(np.median(rb[2:]))
        passes["readback"] = {"kernel": "k_readback", "ms": rb_ms, "bytes_per_pixel": 20,
                              "achieved_gbs": 20.0 * size * size / rb_ms / 1e6, "frac_of_hbm_peak": 20.0 * size * size / rb_ms / 1e6 / peak_gbs}
        passes["sort"] = {"kernels": "k_sort_hist/scan/scatter", "keys": runs, "ms": stages_ms["sort"], "gkeys_per_s": runs / stages_ms["sort"] / 1e6}
        passes["coverage"] = {"kernels": "k_rows/k_rows_long/k_tile_flags", "runs": runs, "ms": stages_ms["coverage"],
                              "gruns_per_s": runs / stages_ms["coverage"] / 1e6}
        # the rows SURVEY 8f marks "next": the PNG the reference's driver writes, encoded on the device from
        # the same framebuffer (20 B per pixel like the readback), and bulk is_point_in_path
        png_size = C.c_size_t(0)
        check(lib.cb200_encode_png(cv, None, 0, C.byref(png_size)))
        png_buf = lib.cb200_host_alloc(png_size.value)
        png = []
        for _ in range(5):
            check(lib.cb200_encode_png(cv, png_buf, png_size.value, None))
            check(lib.cb200_get_stats(cv, C.byref(stats)))
            png.append(stats.png_ms)
        lib.cb200_host_free(png_buf)
        png_ms = float(np.median(png[1:]))
        passes["png_encode"] = {"kernels": "k_png_rows + k_png_finish", "ms": png_ms, "bytes_per_pixel": 20, "file_bytes": png_size.value,
                                "achieved_gbs": 20.0 * size * size / png_ms / 1e6, "frac_of_hbm_peak": 20.0 * size * size / png_ms / 1e6 / peak_gbs}
        from tests.test_hit_testing import scene as hit_scene, _edges as hit_edges
        edges = hit_edges(lib, hit_scene("long_path"))
        pts = np.ascontiguousarray(np.random.default_rng(3).random((1 << 20, 2), dtype=np.float32) * np.float32(256.0))
        inside = np.zeros(len(pts), np.uint8)
        hit_ms, hit = C.c_float(0), []
        for _ in range(4):
            check(lib.cb200_hit_test(cv, edges.ctypes.data, len(edges), pts.ctypes.data, len(pts), inside.ctypes.data, C.byref(hit_ms)))
            hit.append(hit_ms.value)
        passes["hit_test"] = {"kernels": "k_hit_test + k_hit_resolve", "points": len(pts), "edges": len(edges), "ms": float(np.median(hit[1:])),
                              "rule_evaluations_per_s": len(pts) * len(edges) / (float(np.median(hit[1:])) * 1e-3)}
    lib.cb200_canvas_destroy(cv)

    shadow_kw = dict(global_alpha=0.9, shadow_blur=16.0, shadow_color=(0, 0, 0, 0.5))
    if not args.skip_configs and rank == 0:
        # config 3 of BASELINE.json on one GPU: global_alpha 0.9, shadow_blur 16, shadow alpha 0.5
        shadow_frame = H.lower_script(H.tiger_script(size, size, **shadow_kw), size, size)[0]
        cv3 = new_canvas(size, size)
        check(lib.cb200_frame_upload(cv3, C.byref(shadow_frame.frame)))
        rows3 = []
        for i in range(3 + 8):
            check(lib.cb200_frame_replay(cv3, 1))
            check(lib.cb200_get_stats(cv3, C.byref(stats)))
            if i >= 3:
                rows3.append((stats.last_frame_ms, stats.blur_ms, stats.shadow_raster_ms, stats.composite_ms))
        f_ms, b_ms, r_ms, c_ms = [float(np.median(c)) for c in zip(*rows3)]
        plane_px, comp_px = int(stats.shadow_pixels), int(stats.composited_pixels)
        check(lib.cb200_set_stage_timing(cv3, 0))
        spans3 = timed_regions([cv3], 8, 0.3, max_regions=16, all_ranks=False)
        lib.cb200_canvas_destroy(cv3)
        algo3 = 32.0 * comp_px + 20.0 * plane_px              # per-pass bytes of SURVEY 8d for this frame (composite + blur + raster)
        passes["config3_tiger_alpha0.9_shadow_blur16"] = {
            "frames_per_s": 8e3 / float(np.median(spans3)), "frame_ms": float(np.median(spans3)) / 8, "frame_ms_with_stage_events": f_ms,
            "shadow_plane_pixels": plane_px, "composited_pixels": comp_px,
            "frac_of_per_pass_hbm_ceiling": algo3 / (float(np.median(spans3)) / 8 * 1e-3) / 1e9 / peak_gbs,
            # the x sweep rasters the shadow alpha itself (k_blur_x), so the two passes of SURVEY 8d -- alpha raster
            # (4 B per working pixel) and blur (16 B) -- are one measurement; the planes really move 8 B per pixel less
            "blur_and_shadow_raster": {"kernels": "k_blur_x (raster fused), k_blur_y", "ms": b_ms + r_ms, "bytes_per_pixel": 20,
                                       "achieved_gbs": 20.0 * plane_px / (b_ms + r_ms) / 1e6,
                                       "frac_of_hbm_peak": 20.0 * plane_px / (b_ms + r_ms) / 1e6 / peak_gbs,
                                       "blur_only_bytes_per_pixel": 16,
                                       "blur_only_frac_of_hbm_peak": 16.0 * plane_px / (b_ms + r_ms) / 1e6 / peak_gbs},
            "composite": {"kernel": "k_composite<general>", "ms": c_ms, "bytes_per_pixel": 32,
                          "achieved_gbs": 32.0 * comp_px / c_ms / 1e6, "frac_of_hbm_peak": 32.0 * comp_px / c_ms / 1e6 / peak_gbs},
        }
        passes["config4_full_canvas_fills"] = config4_pass(lib, H, _native, args.fill_size, local, peak_gbs)
    if not args.skip_configs:
        passes_5 = config5_pass(lib, H, _native, args.batch, rank, world, local, barrier, torch, dist)
        bands_3 = config3_bands_pass(lib, H, _native, sharding, size, shadow_kw, rank, world, local, barrier, torch, dist)
        if rank == 0:
            passes["config5_batch_of_256x256_canvases"] = passes_5
            if bands_3:
                passes["config3_one_frame_in_scanline_bands"] = bands_3

    # ---- end-to-end arm through the public API with host buffers: `e2e` ----
    # Every frame: fresh canvas state (save/restore + cb200_clear), canvas-script replay on the host
    # (path building + lowering), pinned H2D of the lowered frame, all kernels, get_image_data into
    # page-locked host memory (sRGB/dither kernel + D2H of the 64 MiB RGBA8 image).  `in_flight`
    # canvases are driven from as many host threads (double buffering: one canvas' D2H overlaps the
    # other's kernels); every frame still does all of the above.
    e2e_script = bytes([H.OP["SAVE"]]) + script + bytes([H.OP["RESTORE"]])

    class Lane:
        def __init__(self):
            self.h = lib.cv_create_band(size, size, local, 0, size)
            self.ptr = lib.cb200_host_alloc(size * size * 4)
            self.out = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(size, size, 4))

        def frame(self):
            check(lib.cb200_clear(lib.cv_device(self.h)))
            lib.cv_run_script(self.h, e2e_script, len(e2e_script), None, 0, None)
            lib.cv_get_image_data(self.h, self.ptr, size, size, 4 * size, 0, 0)

        def close(self):
            lib.cv_destroy(self.h)
            self.out = None
            lib.cb200_host_free(self.ptr)

    def run_lanes(lanes, frames_each):
        ts = [threading.Thread(target=lambda l=l: [l.frame() for _ in range(frames_each)]) for l in lanes]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    results = {}
    e2e_frames_each = max(args.steps, 24)
    for n_lanes in (1, args.e2e_lanes):
        lanes = [Lane() for _ in range(n_lanes)]
        run_lanes(lanes, args.warmup)
        if os.environ.get("CB200_E2E_BREAKDOWN") and n_lanes == 1:
            l = lanes[0]
            tc = tr = tg = 0.0
            for _ in range(10):
                a = time.perf_counter(); check(lib.cb200_clear(lib.cv_device(l.h)))
                b = time.perf_counter(); lib.cv_run_script(l.h, e2e_script, len(e2e_script), None, 0, None); lib.cv_flush(l.h)
                c = time.perf_counter(); lib.cv_get_image_data(l.h, l.ptr, size, size, 4 * size, 0, 0)
                d = time.perf_counter()
                tc += b - a; tr += c - b; tg += d - c
            print("e2e breakdown ms: clear %.3f script+lower+submit %.3f get_image_data %.3f" %
                  (tc * 100, tr * 100, tg * 100), file=sys.stderr)
        barrier()
        seconds = run_lanes(lanes, e2e_frames_each)
        barrier()
        results[n_lanes] = (seconds, e2e_frames_each * n_lanes)
        checksum = int(lanes[0].out[::64, ::64].sum())
        [l.close() for l in lanes]
    e2e = {"seconds": results[args.e2e_lanes][0], "frames": results[args.e2e_lanes][1], "serial": results[1], "h2d": frame.upload_bytes,
           "d2h": size * size * 4, "checksum": checksum}

    # ---- max over ranks, aggregate ----
    t_dev = torch.tensor([span_ms / 1e3, e2e["seconds"], min(spans) / 1e3, max(spans) / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    device_s, e2e_s = float(t_dev[0]), float(t_dev[1])
    frames_per_region = args.steps * n_canvases * world
    value = frames_per_region / device_s
    achieved = composited * ALGO_BYTES_PER_COMPOSITED_PIXEL / (comp_ms * 1e-3) / 1e9
    traffic = ncu_traffic("k_composite")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * device_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(size), "canvas": [size, size], "frames_per_step": n_canvases,
                       "parallelism": "frames x%d GPU, %d canvases in flight per GPU (one frame on each per step)" % (world, n_canvases),
                       "replay": "one CUDA graph launch per frame (%d graph replays in all)" % graph_replays if graph_replays
                                 else "stream launches",
                       "l2": "268 MB float framebuffer per frame > 126 MB L2 (inputs larger than L2, no explicit flush)",
                       "composited_pixels_per_frame": composited,
                       "composited_mpix_per_s": composited * value / 1e6,
                       "canvas_mpix_per_s": size * size * value / 1e6},
            "timed": {"regions": len(spans), "steps_per_region": args.steps, "frames_per_region": frames_per_region,
                      "value_from": "median span, max over ranks", "seconds_timed": sum(spans) / 1e3},
            "span_ms": {"median": 1e3 * device_s, "min": 1e3 * float(t_dev[2]), "max": 1e3 * float(t_dev[3]),
                        "rel_spread": (float(t_dev[3]) - float(t_dev[2])) / device_s},
            "roofline": {"bound": "hbm", "kernel": "k_composite", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": composited * ALGO_BYTES_PER_COMPOSITED_PIXEL,
                         "kernel_ms": comp_ms,
                         "dram_frac": (traffic / (comp_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
                         "measured_over": "%d frames on ONE canvas (the kernel alone on the GPU; in the %d-canvas regions its launches "
                                          "overlap other frames' kernels)" % (int(comp_frames.value), n_canvases),
                         "note": "frac uses ALGORITHMIC bytes = 32 B x composited pixels (the contract's figure); the tile compositor "
                                 "keeps overlapping draws in registers and culls occluded ones, so its real DRAM traffic (`traffic`, "
                                 "ncu) is a fraction of that: dram_frac is what the kernel actually asks of HBM"},
            "single_canvas": {"value": single_frames / single_s, "unit": UNIT, "ms_per_frame": 1e3 * single_s / single_frames},
            "stages_ms": stages_ms,
            "gpu_launches": launches_per_region,
            "clocks": clocks.summary(),
        }
        if passes:
            line["passes"] = passes
        line["e2e"] = {"value": e2e["frames"] * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"] * args.e2e_lanes,
                       "d2h_bytes_per_step": e2e["d2h"] * args.e2e_lanes, "frames_per_step": args.e2e_lanes, "in_flight": args.e2e_lanes,
                       "h2d_bytes_per_frame": e2e["h2d"], "d2h_bytes_per_frame": e2e["d2h"],
                       "serial_value": e2e["serial"][1] / e2e["serial"][0],
                       "note": "in_flight canvases, one host thread each, so that one canvas' D2H overlaps the others' kernels; "
                               "serial_value = one canvas, one thread"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(size)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def pin_to_gpu_numa_node(local, world):
    """N ranks x (enqueue thread + e2e threads + 64 MiB pinned buffers per frame) on one NUMA node is what bounded the
    8-GPU end-to-end arm in round 1: keep each rank's threads -- and with them its first-touch pinned allocations -- on
    the CPUs next to its GPU (sysfs via nvidia-smi topo), or on an even slice of the cores when that is unknown."""
    try:
        n = os.cpu_count() or 1
        cpus = None
        out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(local)], capture_output=True, text=True, timeout=10).stdout
        for tok in out.replace(",", " ").split():
            if "-" in tok and tok.replace("-", "").isdigit():
                lo, hi = tok.split("-")
                cpus = (cpus or set()) | set(range(int(lo), int(hi) + 1))
        if world > 1:
            share = max(1, n // world)
            mine = set(range(local * share, min(n, (local + 1) * share)))
            if cpus and len(cpus & mine) >= 2:
                cpus = cpus & mine                          # my slice of the GPU-local CPUs
            elif not cpus or len(cpus) >= n:
                cpus = mine
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def config4_pass(lib, H, _native, size, local, peak_gbs):
    """BASELINE.json config 4: full-canvas fills at 8192^2 -- solid / linear / radial / bicubic draw_image brushes under
    source_copy, exclusive_or and lighter, over a translucent background so that the measured draw blends real pixels.
    Per fill: the compositor's CUDA-event time against 32 B per composited pixel."""
    from canvas_ity_b200.script import ScriptWriter
    out = {"canvas": [size, size], "bytes_per_pixel": 32, "fills": {}}
    cv = C.c_void_p()
    if lib.cb200_canvas_create(size, size, local, C.byref(cv)) != 0:
        return {"error": lib.cb200_last_error().decode()}
    st = _native.Stats()
    bg = ScriptWriter()
    bg.ints("SET_COLOR", 0); bg.raw("4f", 0.9, 0.8, 0.1, 0.6); bg.floats("FILL_RECTANGLE", 0, 0, float(size), float(size))
    bgf = H.lower_script(bg.take(), size, size)[0]
    try:
        assert lib.cb200_submit(cv, C.byref(bgf.frame)) == 0
        for kind in ("solid", "linear", "radial", "image"):
            for opname, op in (("source_copy", 2), ("exclusive_or", 15), ("lighter", 10)):
                frame = H.lower_script(H.config4_script(kind, op, size), size, size)[0]
                assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0, lib.cb200_last_error()
                ts = []
                for i in range(7):
                    assert lib.cb200_frame_replay(cv, 0) == 0
                    assert lib.cb200_get_stats(cv, C.byref(st)) == 0
                    if i >= 2:
                        ts.append(st.composite_ms)
                ms = float(np.median(ts))
                px = int(st.composited_pixels)
                out["fills"]["%s/%s" % (kind, opname)] = {"composite_ms": ms, "composited_pixels": px, "achieved_gbs": 32.0 * px / ms / 1e6,
                                                          "frac_of_hbm_peak": 32.0 * px / ms / 1e6 / peak_gbs}
    finally:
        lib.cb200_canvas_destroy(cv)
    for kind in ("solid", "linear", "radial", "image"):
        out[kind + "_frac_of_hbm_peak"] = float(np.mean([v["frac_of_hbm_peak"] for k, v in out["fills"].items() if k.startswith(kind)]))
    return out


def config5_pass(lib, H, _native, n, rank, world, local, barrier, torch, dist):
    """BASELINE.json config 5: independent 256x256 canvases (8 random fills / strokes + one fill_text each), `n` per GPU
    rendered as ONE device frame; rank r takes canvases i = r (mod world) of the n * world seeded scenes -- no
    collective.  canvases/s from the CUDA-event frame time (max over ranks)."""
    mine = list(range(rank, n * world, world))
    scripts = [H.config5_script(i) for i in mine]
    batch = lib.cv_batch_create(len(mine), 256, 256, local)
    if not batch:
        return {"error": lib.cv_last_error().decode()}
    st = _native.Stats()
    frames_ms, stage = [], None
    try:
        dev = lib.cv_batch_device(batch)
        for rep in range(4):
            assert lib.cb200_clear(dev) == 0
            for i, s in enumerate(scripts):
                lib.cv_run_script(lib.cv_batch_canvas(batch, i), s, len(s), None, 0, None)
            barrier()
            assert lib.cv_batch_flush(batch) == 0, lib.cv_last_error()
            assert lib.cb200_get_stats(dev, C.byref(st)) == 0
            if rep >= 1:
                frames_ms.append(st.last_frame_ms)
                stage = {"geometry": st.geometry_ms, "raster": st.raster_ms, "sort": st.sort_ms, "coverage": st.coverage_ms, "composite": st.composite_ms}
        runs, draws = int(st.raw_runs), int(st.draws)
    finally:
        lib.cv_batch_destroy(batch)
    ms = torch.tensor([float(np.median(frames_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"canvases": n * world, "canvases_per_gpu": n, "frame_ms": float(ms[0]), "canvases_per_s": n * world / (float(ms[0]) * 1e-3),
           "canvases_per_s_per_gpu": n / (float(ms[0]) * 1e-3), "draws": draws, "raw_runs": runs, "stages_ms_rank0": stage,
           "sort_gkeys_per_s": runs / stage["sort"] / 1e6 if stage and stage["sort"] else None}
    out["in_flight"] = config5_in_flight(lib, scripts, n, world, local, barrier, torch, dist)
    return out


def config5_in_flight(lib, scripts, n, world, local, barrier, torch, dist, parts=4, steps=3, regions=3):
    """The same canvases of this rank as `parts` batches on `parts` streams, replayed concurrently from their resident
    frames (cb200_frame_keep + cb200_frame_replay, like the tiger's lanes): most kernels of a batch frame are latency-
    or occupancy-bound, so frames in flight fill one another's idle issue slots.  Span from the earliest begin to the
    latest end event, one untimed frame queued first; canvases/s = canvases x steps / median span (max over ranks)."""
    batches, devs = [], []
    ms = C.c_float()
    try:
        for k in range(parts):
            mine = scripts[k::parts]
            b = lib.cv_batch_create(len(mine), 256, 256, local)
            if not b:
                return {"error": lib.cv_last_error().decode()}
            batches.append(b)
            for i, s in enumerate(mine):
                lib.cv_run_script(lib.cv_batch_canvas(b, i), s, len(s), None, 0, None)
            assert lib.cv_batch_flush(b) == 0, lib.cv_last_error()
            dev = lib.cv_batch_device(b)
            assert lib.cb200_frame_keep(dev) == 0, lib.cb200_last_error()
            assert lib.cb200_set_stage_timing(dev, 0) == 0
            devs.append(dev)
        for d in devs:                                          # first replay verifies the frame, later ones are graph launches
            assert lib.cb200_frame_replay(d, 1) == 0, lib.cb200_last_error()
        for d in devs:
            assert lib.cb200_sync(d) == 0
        spans = []
        for _ in range(regions):
            for d in devs:
                assert lib.cb200_frame_replay(d, 1) == 0
            barrier()
            for d in devs:
                assert lib.cb200_timer_begin(d) == 0
            for _ in range(steps):
                for d in devs:
                    assert lib.cb200_frame_replay(d, 1) == 0
            for d in devs:
                assert lib.cb200_timer_stop(d) == 0
            span = 0.0
            for d in devs:
                assert lib.cb200_timer_end(d, C.byref(ms), None, None) == 0
            for a in devs:
                for b in devs:
                    assert lib.cb200_timer_between(a, b, C.byref(ms)) == 0
                    span = max(span, ms.value)
            spans.append(span)
    finally:
        for b in batches:
            lib.cv_batch_destroy(b)
    span = torch.tensor([float(np.median(spans))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(span, op=dist.ReduceOp.MAX)
    return {"batches_in_flight": parts, "steps": steps, "span_ms": float(span[0]), "frame_ms": float(span[0]) / steps,
            "canvases_per_s": n * world * steps / (float(span[0]) * 1e-3), "canvases_per_s_per_gpu": n * steps / (float(span[0]) * 1e-3)}


def config3_bands_pass(lib, H, _native, sharding, size, shadow_kw, rank, world, local, barrier, torch, dist):
    """BASELINE.json config 3 as ONE frame: rank r renders scanline band r of the 4096^2 shadowed tiger
    (cb200_canvas_create_band), converts it to RGBA8 straight into an NCCL buffer (cb200_read_rgba8_into) and the bands are
    all_gathered into one image on every rank.  Everything is queued on the canvas' own stream (the gather too: torch is
    pointed at it), no host synchronisation per frame; CUDA events around K frames, max over ranks.  Strong scaling."""
    if world == 1:
        return None
    y0, rows = sharding.band(size, rank, world)
    max_rows = (size + world - 1) // world
    frame = H.lower_script(H.tiger_script(size, size, **shadow_kw), size, size)[0]
    cv = C.c_void_p()
    assert lib.cb200_canvas_create_band(size, size, y0, rows, local, C.byref(cv)) == 0, lib.cb200_last_error()
    stream = torch.cuda.ExternalStream(lib.cb200_stream(cv), device=torch.device("cuda", local))
    band = torch.zeros((max_rows, size, 4), dtype=torch.uint8, device="cuda")
    gathered = torch.empty((world * max_rows, size, 4), dtype=torch.uint8, device="cuda")
    assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
    assert lib.cb200_set_stage_timing(cv, 0) == 0
    ms, gather_ms = C.c_float(), []
    K = 16

    def one_frame(with_gather=True):
        assert lib.cb200_frame_replay(cv, 1) == 0, lib.cb200_last_error()
        assert lib.cb200_read_rgba8_into(cv, C.c_void_p(band.data_ptr()), size, rows, 0, y0) == 0
        if with_gather:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered.view(-1), band.view(-1))

    spans, render_spans = [], []
    for with_gather, dest in ((True, spans), (False, render_spans)):
        for rep in range(4):
            one_frame(with_gather)
            barrier()
            assert lib.cb200_timer_begin(cv) == 0
            for _ in range(K):
                one_frame(with_gather)
            assert lib.cb200_timer_end(cv, C.byref(ms), None, None) == 0
            barrier()
            if rep:
                dest.append(ms.value / K)
    st = _native.Stats()
    assert lib.cb200_set_stage_timing(cv, 1) == 0
    for _ in range(3):
        assert lib.cb200_frame_replay(cv, 1) == 0
        assert lib.cb200_get_stats(cv, C.byref(st)) == 0
    stage = {"geometry": st.geometry_ms, "raster": st.raster_ms, "sort": st.sort_ms, "coverage": st.coverage_ms,
             "shadow_raster": st.shadow_raster_ms, "blur": st.blur_ms, "composite": st.composite_ms}
    checksum = int(gathered[::97, ::89].sum().item())
    lib.cb200_canvas_destroy(cv)
    t = torch.tensor([float(np.median(spans)), float(np.median(render_spans))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frame_ms, render_ms = float(t[0]), float(t[1])
    return {"frames_per_s": 1e3 / frame_ms, "frame_ms": frame_ms, "render_and_convert_ms": render_ms,
            "all_gather_ms": max(0.0, frame_ms - render_ms), "gathered_bytes": world * max_rows * size * 4, "bands": world,
            "stages_ms_rank0": stage, "checksum": checksum,
            "note": "one 4096^2 frame sharded by scanline bands (shadow planes clipped to band +- blur halo), RGBA8 bands "
                    "all_gathered over NCCL on the canvas stream; compare with config3 frame_ms on one GPU"}


if __name__ == "__main__":
    main()
