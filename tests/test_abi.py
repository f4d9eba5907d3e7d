"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from tests import harness as H
from canvas_ity_b200 import _native


def declared_functions(header):
    text = open(os.path.join(H.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cb200|cv)_[a-z0-9_]+)\s*\(", text)) - {"cv_frame_fn", "cv_read_fn", "cv_write_fn"})


@pytest.mark.parametrize("header", ["canvas_b200.h", "canvas_b200_api.h"])
def test_every_declared_symbol_is_exported(header):
    lib = C.CDLL(_native.LIB_PATH)
    names = declared_functions(header)
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), "%s declared in include/%s but not exported" % (name, header)
        assert name in _native.SIGNATURES, "%s has no ctypes signature in _native.py" % name


def test_abi_version_and_backend_name():
    lib = _native.load()
    assert lib.cb200_abi_version() == 2
    assert lib.cv_backend_name() == b"b200"


def test_no_device_fails_loudly():
    """Without a CUDA device there is no fallback: creation fails with an error string."""
    lib = _native.load()
    if lib.cb200_device_count() > 0:
        pytest.skip("a device is present")
    out = C.c_void_p()
    rc = lib.cb200_canvas_create(16, 16, 0, C.byref(out))
    assert rc == -1 and not out.value
    assert b"no CUDA device" in lib.cb200_last_error()
    assert not lib.cv_create(16, 16)
    assert b"CUDA" in lib.cv_last_error()
    import canvas_ity_b200
    with pytest.raises(RuntimeError):
        canvas_ity_b200.Canvas(16, 16)


def test_bad_arguments_are_rejected():
    lib = _native.load()
    out = C.c_void_p()
    assert lib.cb200_canvas_create(0, 16, 0, C.byref(out)) == -2
    assert lib.cb200_canvas_create(16, 40000, 0, C.byref(out)) == -2
    assert lib.cb200_canvas_create_band(16, 16, 8, 16, 0, C.byref(out)) == -2
    assert lib.cb200_submit(None, None) == -2


def test_front_header_is_drop_in_for_reference_drivers():
    """The reference's own tiger demo must compile unmodified against include/canvas_ity.hpp
    (only where /root/reference is mounted; the include line is redirected, nothing is copied)."""
    src = "/root/reference/demos/tiger/tiger.cpp"
    if not os.path.exists(src):
        pytest.skip("reference not mounted here")
    import subprocess, tempfile
    tmp = tempfile.mkdtemp(prefix="cb200_dropin_")
    os.makedirs(os.path.join(tmp, "src"))
    os.makedirs(os.path.join(tmp, "demos", "tiger"))
    os.symlink(os.path.join(H.ROOT, "include", "canvas_ity.hpp"), os.path.join(tmp, "src", "canvas_ity.hpp"))
    os.symlink(src, os.path.join(tmp, "demos", "tiger", "tiger.cpp"))
    obj = os.path.join(tmp, "tiger.o")
    subprocess.check_call(["g++", "-std=c++17", "-O0", "-c", os.path.join(tmp, "demos", "tiger", "tiger.cpp"), "-o", obj])
    exe = os.path.join(tmp, "tiger")
    subprocess.check_call(["g++", obj, "-o", exe, "-L" + os.path.dirname(_native.LIB_PATH), "-lcanvas_b200",
                           "-Wl,-rpath," + os.path.dirname(_native.LIB_PATH)])
    assert os.path.exists(exe)


def test_record_sizes_match_the_python_mirror():
    lib = _native.load()
    assert lib.cb200_struct_size(0) == _native.SIZEOF_DRAW
    assert lib.cb200_struct_size(1) == _native.SIZEOF_SUBPATH
    assert lib.cb200_struct_size(2) == _native.SIZEOF_BRUSH
    assert lib.cb200_struct_size(3) == _native.SIZEOF_IMAGE
    assert lib.cb200_struct_size(4) == C.sizeof(_native.Frame)
    assert lib.cb200_struct_size(5) == _native.SIZEOF_GLYPH_SEG
    assert lib.cb200_struct_size(6) == _native.SIZEOF_GLYPH_OUTLINE
    assert lib.cb200_struct_size(7) == _native.SIZEOF_GLYPH_ATLAS
    assert lib.cb200_struct_size(8) == _native.SIZEOF_GLYPH_INST


def test_lowering_of_the_tiger_is_one_frame_of_305_draws():
    frames = H.lower_script(H.tiger_script(512, 512), 512, 512)
    assert len(frames) == 1 and frames[0].n_draws == 305
