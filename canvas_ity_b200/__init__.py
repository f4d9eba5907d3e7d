"""canvas_ity_b200 -- B200-native rasterise + composite back end behind the canvas_ity API.

Layout: ``csrc/`` holds the sm_100a kernels and the C++ front end (built into
``libcanvas_b200.so``); this package is the ctypes mirror of the reference's
``canvas_ity::canvas`` class plus multi-GPU helpers.  The CUDA library is
mandatory: importing works without it, creating a canvas does not.
"""
from .canvas import *          # noqa: F401,F403  (enum names mirror the reference namespace)
from .canvas import Canvas
from . import _native, script

__all__ = ["Canvas", "_native", "script"]
