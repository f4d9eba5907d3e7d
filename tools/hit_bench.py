#!/usr/bin/env python
"""Bulk hit testing: cb200_hit_test against the reference's is_point_in_path loop.

Workload: the path of tests/test_hit_testing.py::scene("long_path") (40 closed subpaths x 30 cubics), N
random query points.  Reports the kernel time (CUDA events), the end-to-end call (host buffers, H2D + kernels
+ D2H), rule evaluations per second, and the reference (one is_point_in_path call per point, which
re-flattens the path every call, hpp:3105) timed on a bounded sample on one host core."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from tests.test_hit_testing import scene, _edges

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
lib = H.product_library()
script = scene("long_path")
edges = _edges(lib, script)
rng = np.random.default_rng(3)
q = np.ascontiguousarray(rng.random((N, 2), dtype=np.float32) * np.float32(256.0))
out = {"workload": "%d points x %d edges" % (N, len(edges))}
if lib.cb200_device_count() > 0:
    h = lib.cv_create(256, 256)
    H._run(lib, h, script)
    dev = lib.cv_device(h)
    got = np.zeros(N, np.uint8)
    ms = C.c_float(0)
    kern, e2e = [], []
    for rep in range(6):
        t0 = time.perf_counter()
        assert lib.cb200_hit_test(dev, edges.ctypes.data, len(edges), q.ctypes.data, N, got.ctypes.data, C.byref(ms)) == 0
        e2e.append(time.perf_counter() - t0)
        kern.append(ms.value)
    out["gpu"] = {"kernel_ms": min(kern[1:]), "e2e_ms": min(e2e[1:]) * 1e3,
                  "rule_evaluations_per_s": N * len(edges) / (min(kern[1:]) * 1e-3),
                  "points_per_s_e2e": N / min(e2e[1:]), "inside_fraction": float(got.mean())}
    lib.cv_destroy(h)
ref = H.reference_library(fast=True) or H.reference_library()
if ref is not None:
    sample = min(N, 2000)
    h = ref.cv_create(256, 256)
    H._run(ref, h, script)
    want = np.zeros(sample, np.uint8)
    t0 = time.perf_counter()
    ref.cv_points_in_path(h, q.ctypes.data, sample, want.ctypes.data)
    dt = time.perf_counter() - t0
    ref.cv_destroy(h)
    out["reference_cpu"] = {"points_per_s": sample / dt, "sample": "%d points, one host core" % sample}
    if "gpu" in out:
        out["identical_on_sample"] = bool(np.array_equal(got[:sample], want))
print(json.dumps(out, indent=1))
