"""ctypes loader for libcanvas_b200.so (the product) -- fails loudly, never falls back.

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C canvas_ity_b200/csrc``.  Signatures mirror include/canvas_b200.h and
include/canvas_b200_api.h one to one.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcanvas_b200.so")


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("draws", "cubics", "line_points", "edges", "raw_runs", "tile_entries",
                                           "composited_pixels", "shadow_pixels", "kernel_launches")] + \
               [(n, C.c_float) for n in ("last_frame_ms", "composite_ms", "raster_ms", "sort_ms", "geometry_ms",
                                         "readback_ms", "coverage_ms", "shadow_raster_ms", "blur_ms", "png_ms")] + \
               [("graph_replays", C.c_uint32)]


class Frame(C.Structure):          # cb200_frame
    _fields_ = [("draws", C.c_void_p), ("n_draws", C.c_uint32),
                ("subpaths", C.c_void_p), ("n_subpaths", C.c_uint32),
                ("points", C.c_void_p), ("n_points", C.c_uint32),
                ("brushes", C.c_void_p), ("n_brushes", C.c_uint32),
                ("colors", C.c_void_p), ("stops", C.c_void_p), ("n_colors", C.c_uint32),
                ("dashes", C.c_void_p), ("n_dashes", C.c_uint32),
                ("images", C.c_void_p), ("n_images", C.c_uint32),
                ("texels", C.c_void_p), ("texel_bytes", C.c_uint64),
                ("atlases", C.c_void_p), ("n_atlases", C.c_uint32),
                ("glyphs", C.c_void_p), ("n_glyphs", C.c_uint32), ("n_glyph_points", C.c_uint32)]


FRAME_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(Frame))
SIZEOF_DRAW, SIZEOF_SUBPATH, SIZEOF_BRUSH, SIZEOF_IMAGE = 140, 16, 48, 16
SIZEOF_GLYPH_SEG, SIZEOF_GLYPH_OUTLINE, SIZEOF_GLYPH_ATLAS, SIZEOF_GLYPH_INST = 16, 24, 56, 36


class OwnedFrame:
    """Deep copy of a cb200_frame (the pointers handed to a frame tap die with the callback)."""

    def __init__(self, f):
        def grab(ptr, nbytes):
            return C.create_string_buffer(C.string_at(ptr, nbytes), nbytes) if ptr and nbytes else C.create_string_buffer(1)
        self.parts = {
            "draws": grab(f.draws, f.n_draws * SIZEOF_DRAW), "subpaths": grab(f.subpaths, f.n_subpaths * SIZEOF_SUBPATH),
            "points": grab(f.points, f.n_points * 8), "brushes": grab(f.brushes, f.n_brushes * SIZEOF_BRUSH),
            "colors": grab(f.colors, f.n_colors * 16), "stops": grab(f.stops, f.n_colors * 4),
            "dashes": grab(f.dashes, f.n_dashes * 4), "images": grab(f.images, f.n_images * SIZEOF_IMAGE),
            "texels": grab(f.texels, f.texel_bytes),
            # atlas snapshots are copied by value; the arrays they point at live as long as the process
            "atlases": grab(f.atlases, f.n_atlases * SIZEOF_GLYPH_ATLAS),
            "glyphs": grab(f.glyphs, f.n_glyphs * SIZEOF_GLYPH_INST)}
        self.frame = Frame()
        for name, _ in Frame._fields_:
            if name in self.parts:
                setattr(self.frame, name, C.cast(self.parts[name], C.c_void_p))
            else:
                setattr(self.frame, name, getattr(f, name))
        self.n_draws = f.n_draws
        self.n_points = f.n_points
        self.n_glyphs = f.n_glyphs
        self.n_glyph_points = f.n_glyph_points
        self.upload_bytes = sum(len(p) for p in self.parts.values())

# name -> (restype, argtypes); every symbol the two headers declare
SIGNATURES = {
    # include/canvas_b200.h
    "cb200_canvas_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cb200_canvas_create_band": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_void_p)]),
    "cb200_canvas_destroy": (None, [C.c_void_p]),
    "cb200_batch_create": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_void_p)]),
    "cb200_batch_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "cb200_batch_read_rgba8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p] + [C.c_int] * 5),
    "cb200_batch_read_f32": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "cb200_batch_write_rgba8": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p] + [C.c_int] * 5),
    "cb200_batch_masks_keep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "cb200_submit": (C.c_int, [C.c_void_p, C.POINTER(Frame)]),
    "cb200_frame_upload": (C.c_int, [C.c_void_p, C.POINTER(Frame)]),
    "cb200_frame_replay": (C.c_int, [C.c_void_p, C.c_int]),
    "cb200_frame_keep": (C.c_int, [C.c_void_p]),
    "cb200_set_graph_replay": (C.c_int, [C.c_void_p, C.c_int]),
    "cb200_sync": (C.c_int, [C.c_void_p]),
    "cb200_read_rgba8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5),
    "cb200_write_rgba8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5),
    "cb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "cb200_host_free": (None, [C.c_void_p]),
    "cb200_read_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cb200_read_mask": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "cb200_clear": (C.c_int, [C.c_void_p]),
    "cb200_masks_keep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "cb200_hit_test": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_float)]),
    "cb200_read_bgra8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb200_framebuffer_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cb200_read_rgba8_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cb200_encode_png": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "cb200_read_rgba8_into": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4),
    "cb200_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "cb200_debug_lines": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "cb200_debug_runs": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "cb200_debug_shadow_box": (None, [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "cb200_debug_loop_runs": (C.c_int64, [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "cb200_debug_join_math": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]),
    "cb200_last_error": (C.c_char_p, []),
    "cb200_set_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "cb200_timer_begin": (C.c_int, [C.c_void_p]),
    "cb200_timer_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "cb200_stream": (C.c_void_p, [C.c_void_p]),
    "cb200_timer_stop": (C.c_int, [C.c_void_p]),
    "cb200_timer_between": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "cb200_abi_version": (C.c_int, []),
    "cb200_device_count": (C.c_int, []),
    "cb200_struct_size": (C.c_int, [C.c_int]),
    # include/canvas_b200_api.h
    "cv_create": (C.c_void_p, [C.c_int, C.c_int]),
    "cv_create_band": (C.c_void_p, [C.c_int] * 5),
    "cv_create_tapped": (C.c_void_p, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cv_destroy": (None, [C.c_void_p]),
    "cv_run_script": (C.c_long, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "cv_get_image_data": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5),
    "cv_put_image_data": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5),
    "cv_is_point_in_path": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "cv_measure_text": (C.c_float, [C.c_void_p, C.c_char_p]),
    "cv_points_in_path": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cv_path_edges": (C.c_long, [C.c_void_p, C.c_void_p, C.c_long]),
    "cv_flush": (C.c_int, [C.c_void_p]),
    "cv_set_text_instancing": (C.c_int, [C.c_void_p, C.c_int]),
    "cv_write_tga": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cv_write_png": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cv_batch_create": (C.c_void_p, [C.c_int] * 4),
    "cv_batch_canvas": (C.c_void_p, [C.c_void_p, C.c_int]),
    "cv_batch_flush": (C.c_int, [C.c_void_p]),
    "cv_batch_get_image_data": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5),
    "cv_batch_read_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cv_batch_device": (C.c_void_p, [C.c_void_p]),
    "cv_batch_destroy": (None, [C.c_void_p]),
    "cv_read_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cv_device": (C.c_void_p, [C.c_void_p]),
    "cv_last_error": (C.c_char_p, []),
    "cv_backend_name": (C.c_char_p, []),
}


def bind(lib, names=None):
    for name, (res, args) in SIGNATURES.items():
        if names is not None and name not in names:
            continue
        fn = getattr(lib, name)       # AttributeError if the symbol is missing: loud on purpose
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """Load the CUDA back end.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "canvas_ity_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib
