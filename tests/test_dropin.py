"""The drop-in claim, checked by the reference itself: test/test.cpp, UNMODIFIED, compiled against
include/canvas_ity.hpp and linked with libcanvas_b200.so (oracle/Makefile target _ref/test_dropin, built where
/root/reference is mounted; the binary travels to the GPU box).  The driver renders its 76 cases through the
public canvas API, reads them back with get_image_data and compares each image hash with its own table
(test.cpp:2186-2261), Hamming distance <= 5 (test.cpp:2618)."""
import os
import re
import subprocess

import pytest

from tests import harness as H

BINARY = os.path.join(H.ROOT, "oracle", "_ref", "test_dropin")


def test_reference_test_driver_builds_against_the_drop_in_header():
    if not os.path.exists("/root/reference/test/test.cpp"):
        pytest.skip("reference not mounted here")
    subprocess.check_call(["make", "-C", os.path.join(H.ROOT, "oracle"), "_ref/test_dropin"], stdout=subprocess.DEVNULL)
    assert os.access(BINARY, os.X_OK)
    if H.product_library().cb200_device_count() == 0:
        # no CPU fallback: without a device the very first canvas constructor fails loudly
        out = subprocess.run([BINARY, "--plain", "--subset", "scale_uniform"], capture_output=True, text=True)
        assert out.returncode != 0 and "no CUDA device" in out.stderr


@pytest.mark.gpu
def test_reference_test_driver_passes_its_own_76_hashes_on_the_gpu():
    if H.product_library().cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    if not os.access(BINARY, os.X_OK):
        pytest.skip("oracle/_ref/test_dropin was not built (needs /root/reference at build time)")
    out = subprocess.run([BINARY, "--plain"], capture_output=True, text=True, timeout=600)
    lines = out.stdout.splitlines()
    passed = [l for l in lines if "[PASS]" in l]
    failed = [l for l in lines if "[FAIL]" in l]
    assert not failed, "\n".join(failed)
    assert len(passed) == 76, out.stdout[-2000:] + out.stderr[-2000:]
    assert re.search(r"^0 failed,", lines[-1]) and out.returncode == 0


@pytest.mark.gpu
def test_tiger_demo_unmodified_writes_the_same_tga(tmp_path):
    """demos/tiger/tiger.cpp, unmodified: built against the drop-in header it renders on the GPU and writes
    tiger.tga (tiger.cpp:4333-4345); the same source built against the reference's header is the CPU render."""
    import numpy as np
    if H.product_library().cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    ours, ref = (os.path.join(H.ROOT, "oracle", "_ref", n) for n in ("tiger_dropin", "tiger_ref"))
    if not (os.access(ours, os.X_OK) and os.access(ref, os.X_OK)):
        pytest.skip("oracle/_ref tiger binaries were not built (need /root/reference at build time)")
    files = {}
    for name, exe in (("ours", ours), ("ref", ref)):
        d = tmp_path / name
        d.mkdir()
        out = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "Best:" in out.stdout, out.stderr[-1000:]
        files[name] = (d / "tiger.tga").read_bytes()
    assert len(files["ours"]) == len(files["ref"]) == 18 + 733 * 757 * 4
    assert files["ours"][:18] == files["ref"][:18]
    a = np.frombuffer(files["ours"][18:], np.uint8).reshape(757, 733, 4)[..., [2, 1, 0, 3]]    # BGRA -> RGBA
    b = np.frombuffer(files["ref"][18:], np.uint8).reshape(757, 733, 4)[..., [2, 1, 0, 3]]
    da, dc, n_off = H.rgba8_mismatch(a, b)
    assert n_off == 0, "alpha %d colour %.2f, %d pixels beyond 1 LSB" % (da, dc, n_off)
