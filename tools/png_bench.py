#!/usr/bin/env python
"""PNG off the device: cb200_encode_png against get_image_data + the reference driver's
write_png (test/test.cpp:2415-2507) on the CPU.  Tiger at SIZE x SIZE (default 4096)."""
import ctypes as C, json, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import harness as H
from canvas_ity_b200 import _native

SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = H.product_library()
out = {"workload": "tiger %dx%d -> PNG file image (%d bytes)" % (SIZE, SIZE, 76 + SIZE * (6 + 4 * SIZE))}
cv = lib.cv_create(SIZE, SIZE)
H._run(lib, cv, H.tiger_script(SIZE, SIZE))
lib.cv_flush(cv)                                   # the draws are queued in the front end until something reads
dev = lib.cv_device(cv)
n = 76 + SIZE * (6 + 4 * SIZE)
buf = lib.cb200_host_alloc(n)
kern, e2e = [], []
for rep in range(6):
    t0 = time.perf_counter()
    assert lib.cb200_encode_png(dev, buf, n, None) == 0
    e2e.append(time.perf_counter() - t0)
    st = _native.Stats()
    lib.cb200_get_stats(dev, C.byref(st))
    kern.append(st.png_ms)
png = C.string_at(buf, n)
pixels = SIZE * SIZE
out["gpu"] = {"kernel_ms": min(kern[1:]), "algorithmic_bytes_per_pixel": 20,
              "achieved_gbs": pixels * 20 / (min(kern[1:]) * 1e-3) / 1e9,
              "e2e_ms_pinned_host_buffer": min(e2e[1:]) * 1e3}
img = np.zeros((SIZE, SIZE, 4), np.uint8)
t = []
for rep in range(3):
    t0 = time.perf_counter()
    lib.cv_get_image_data(cv, img.ctypes.data, SIZE, SIZE, 4 * SIZE, 0, 0)
    t.append(time.perf_counter() - t0)
out["gpu"]["get_image_data_ms_for_comparison"] = min(t) * 1e3
ref = H.reference_library(fast=True) or H.reference_library()
if ref is not None:
    fn = ref.ref_write_png
    fn.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]; fn.restype = None
    path = os.path.join(tempfile.mkdtemp(), "ref.png")
    t = []
    for rep in range(3):
        t0 = time.perf_counter()
        fn(path.encode(), img.ctypes.data, SIZE, SIZE)
        t.append(time.perf_counter() - t0)
    out["reference_cpu"] = {"write_png_ms": min(t) * 1e3, "note": "the reference driver's write_png over the same pixels, one host core, file in a temp dir"}
    out["identical_files"] = open(path, "rb").read() == png
lib.cb200_host_free(buf)
lib.cv_destroy(cv)
print(json.dumps(out, indent=1))
