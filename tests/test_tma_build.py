"""The bulk-asynchronous-copy (TMA engine) build of the lean compositor -- CB200_TMA=1, k_composite<0, *, true>:
framebuffer rows travel as cp.async.bulk copies through shared memory instead of 16-byte LDG/STG per lane -- does the
same arithmetic on the same pixels, so its framebuffer must equal the default build's bit for bit.  The switch is read
once per process, hence the subprocess.  (It is the slower of the two builds and not the default: DESIGN.md K7,
profiles/r02_tma_ab_compositor.txt.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import harness as H

pytestmark = pytest.mark.gpu

RENDER = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, {root!r})
from tests import harness as H
lib = H.product_library()
out = []
# fresh canvases (bulk stores only), one of them with partial tiles at the right and bottom edges ...
for size in (512, 733):
    out.append(H.render_script(lib, H.tiger_script(size, size), size, size)["f32"].ravel())
# ... and a second frame over the pixels of the first (bulk loads of the old pixels wherever no opaque draw covers the tile)
size = 320
frame = H.lower_script(H.tiger_script(size, size, global_alpha=0.5), size, size)[0]
cv = C.c_void_p()
assert lib.cb200_canvas_create(size, size, 0, C.byref(cv)) == 0
for _ in range(2):
    assert lib.cb200_submit(cv, C.byref(frame.frame)) == 0, lib.cb200_last_error()
twice = np.zeros((size, size, 4), np.float32)
assert lib.cb200_read_f32(cv, twice.ctypes.data) == 0
lib.cb200_canvas_destroy(cv)
out.append(twice.ravel())
np.save({path!r}, np.concatenate(out))
"""


def _render(tmp_path, name, tma):
    path = str(tmp_path / (name + ".npy"))
    env = dict(os.environ, CB200_TMA="1" if tma else "0")
    subprocess.run([sys.executable, "-c", RENDER.format(root=H.ROOT, path=path)], check=True, env=env, timeout=300)
    return np.load(path)


def test_bulk_copy_build_equals_the_default_build(tmp_path):
    lib = H.product_library()
    if lib.cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    plain, bulk = _render(tmp_path, "plain", False), _render(tmp_path, "bulk", True)
    assert float(np.abs(plain).sum()) > 0.0
    assert np.array_equal(plain.view(np.uint32), bulk.view(np.uint32))
