"""Python mirror of ``canvas_ity::canvas`` (reference src/canvas_ity.hpp:177-1148).

Same method names, argument order and silent-no-op behaviour as the reference's
public API; public data members are attributes.  Calls are encoded into a canvas
script and replayed by the C++ front end (libcanvas_b200.so), which lowers draws
and runs them on the GPU.  Synchronous calls (get_image_data, is_point_in_path,
measure_text, set_font) flush first.  No CPU rendering happens here.
"""
import ctypes as C

import numpy as np

from . import _native
from .script import ScriptWriter

# enums, reference :146-156
source_in, source_copy, source_out, destination_in = 1, 2, 3, 4
destination_atop, lighter, destination_over, destination_out = 7, 10, 11, 12
source_atop, source_over, exclusive_or = 13, 14, 15
butt, square, circle = 0, 1, 2
miter, bevel, rounded = 0, 1, 2
fill_style, stroke_style = 0, 1
repeat, repeat_x, repeat_y, no_repeat = 0, 1, 2, 3
leftward, rightward, center, start, ending = 0, 1, 2, 0, 1
alphabetic, top, middle, bottom, hanging, ideographic = 0, 1, 2, 3, 4, 3

_FIELDS = {"global_composite_operation": ("SET_COMPOSITE", "i", source_over),
           "shadow_offset_x": ("SET_SHADOW_OFFSET_X", "f", 0.0), "shadow_offset_y": ("SET_SHADOW_OFFSET_Y", "f", 0.0),
           "line_cap": ("SET_LINE_CAP", "i", butt), "line_join": ("SET_LINE_JOIN", "i", miter),
           "line_dash_offset": ("SET_LINE_DASH_OFFSET", "f", 0.0),
           "text_align": ("SET_TEXT_ALIGN", "i", start), "text_baseline": ("SET_TEXT_BASELINE", "i", alphabetic)}


def _image_bytes(image, width, height, stride):
    if image is None:
        return None
    data = image.tobytes() if hasattr(image, "tobytes") else bytes(image)
    if width <= 0 or height <= 0:
        return data[:1] or b"\0"             # a no-op in the reference (hpp:2847, 3323, 3392): nothing is read
    if stride < 0:
        raise ValueError("canvas_ity_b200: a negative stride cannot be expressed for a Python buffer")
    need = (height - 1) * stride + width * 4
    if len(data) < need:
        raise ValueError("canvas_ity_b200: image buffer holds %d bytes, %d x %d pixels with stride %d need %d"
                         % (len(data), width, height, stride, need))
    return data[:need]


class Canvas:
    def __init__(self, width, height, library=None, handle=None, device=None, band=None):
        self._lib = library or _native.load()
        self.width, self.height = int(width), int(height)
        if handle is not None:
            self._h = handle
        elif band is not None:
            self._h = self._lib.cv_create_band(self.width, self.height, int(device or 0), int(band[0]), int(band[1]))
        else:
            self._h = self._lib.cv_create(self.width, self.height)
        if not self._h:
            raise RuntimeError("canvas_ity_b200: " + self._lib.cv_last_error().decode())
        self._band = None if band is None else (int(band[0]), int(band[1]))
        self._device_index = int(device or 0)
        self._w = ScriptWriter()
        self._state = {k: v[2] for k, v in _FIELDS.items()}
        self._stack = []
        self.queries = []

    # -- public data members (sampled when the next call is recorded, like the reference) --
    def __getattr__(self, name):
        if name in _FIELDS:
            return self.__dict__["_state"][name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _FIELDS:
            op, kind, _ = _FIELDS[name]
            self._state[name] = value
            (self._w.ints if kind == "i" else self._w.floats)(op, value)
        else:
            object.__setattr__(self, name, value)

    # -- plumbing --
    def flush(self):
        script = self._w.take()
        if script:
            q = (C.c_uint32 * (4 * 256))()
            nq = C.c_int(0)
            n = self._lib.cv_run_script(self._h, script, len(script), q, 256, C.byref(nq))
            if n < 0:
                raise RuntimeError("malformed canvas script")
            self.queries += [tuple(q[i * 4:i * 4 + 3]) for i in range(min(nq.value, 256))]
        self._lib.cv_flush(self._h)

    def run_script(self, script):
        """Replay a recorded canvas script (tests/golden/scripts/*.cvs)."""
        self.flush()
        q = (C.c_uint32 * (4 * 4096))()
        nq = C.c_int(0)
        n = self._lib.cv_run_script(self._h, script, len(script), q, 4096, C.byref(nq))
        if n < 0:
            raise RuntimeError("malformed canvas script")
        self.queries += [tuple(q[i * 4:i * 4 + 3]) for i in range(min(nq.value, 4096))]
        return n

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- transforms --
    def scale(self, x, y): self._w.floats("SCALE", x, y)
    def rotate(self, angle): self._w.floats("ROTATE", angle)
    def translate(self, x, y): self._w.floats("TRANSLATE", x, y)
    def transform(self, a, b, c, d, e, f): self._w.floats("TRANSFORM", a, b, c, d, e, f)
    def set_transform(self, a, b, c, d, e, f): self._w.floats("SET_TRANSFORM", a, b, c, d, e, f)

    # -- compositing, shadows, line styles --
    def set_global_alpha(self, alpha): self._w.floats("SET_GLOBAL_ALPHA", alpha)
    def set_shadow_color(self, r, g, b, a): self._w.floats("SET_SHADOW_COLOR", r, g, b, a)
    def set_shadow_blur(self, level): self._w.floats("SET_SHADOW_BLUR", level)
    def set_line_width(self, width): self._w.floats("SET_LINE_WIDTH", width)
    def set_miter_limit(self, limit): self._w.floats("SET_MITER_LIMIT", limit)

    def set_line_dash(self, segments, count=None):
        if segments is None:
            self._w.ints("SET_LINE_DASH_NULL", 0 if count is None else count)
            return
        count = len(segments) if count is None else count
        self._w.ints("SET_LINE_DASH", count)
        self._w.raw("%df" % count, *segments[:count])

    # -- brushes --
    def set_color(self, which, r, g, b, a):
        self._w.ints("SET_COLOR", which); self._w.raw("4f", r, g, b, a)

    def set_linear_gradient(self, which, sx, sy, ex, ey):
        self._w.ints("SET_LINEAR_GRADIENT", which); self._w.raw("4f", sx, sy, ex, ey)

    def set_radial_gradient(self, which, sx, sy, sr, ex, ey, er):
        self._w.ints("SET_RADIAL_GRADIENT", which); self._w.raw("6f", sx, sy, sr, ex, ey, er)

    def add_color_stop(self, which, offset, r, g, b, a):
        self._w.ints("ADD_COLOR_STOP", which); self._w.raw("5f", offset, r, g, b, a)

    def set_pattern(self, which, image, width, height, stride, repetition):
        self._w.ints("SET_PATTERN", which, width, height, stride, repetition)
        self._w.blob(_image_bytes(image, width, height, stride))

    # -- paths --
    def begin_path(self): self._w.bare("BEGIN_PATH")
    def move_to(self, x, y): self._w.floats("MOVE_TO", x, y)
    def close_path(self): self._w.bare("CLOSE_PATH")
    def line_to(self, x, y): self._w.floats("LINE_TO", x, y)
    def quadratic_curve_to(self, cx, cy, x, y): self._w.floats("QUADRATIC_CURVE_TO", cx, cy, x, y)
    def bezier_curve_to(self, c1x, c1y, c2x, c2y, x, y): self._w.floats("BEZIER_CURVE_TO", c1x, c1y, c2x, c2y, x, y)
    def arc_to(self, vx, vy, x, y, radius): self._w.floats("ARC_TO", vx, vy, x, y, radius)

    def arc(self, x, y, radius, start_angle, end_angle, counter_clockwise=False):
        self._w.floats("ARC", x, y, radius, start_angle, end_angle); self._w.raw("i", 1 if counter_clockwise else 0)

    def rectangle(self, x, y, w, h): self._w.floats("RECTANGLE", x, y, w, h)

    # -- drawing --
    def fill(self): self._w.bare("FILL")
    def stroke(self): self._w.bare("STROKE")
    def clip(self): self._w.bare("CLIP")
    def clear_rectangle(self, x, y, w, h): self._w.floats("CLEAR_RECTANGLE", x, y, w, h)
    def fill_rectangle(self, x, y, w, h): self._w.floats("FILL_RECTANGLE", x, y, w, h)
    def stroke_rectangle(self, x, y, w, h): self._w.floats("STROKE_RECTANGLE", x, y, w, h)

    def is_point_in_path(self, x, y):
        self.flush()
        return bool(self._lib.cv_is_point_in_path(self._h, x, y))

    def points_in_path(self, xy):
        """is_point_in_path (hpp:3101-3132) for many device-space points at once: `xy` is (n, 2) float32,
        the result (n,) bool -- one device launch (cb200_hit_test) instead of n host walks of the path."""
        import numpy as np
        self.flush()
        q = np.ascontiguousarray(np.asarray(xy, np.float32).reshape(-1, 2))
        inside = np.zeros(len(q), np.uint8)
        if self._lib.cv_points_in_path(self._h, q.ctypes.data, len(q), inside.ctypes.data) != 0:
            raise RuntimeError(self._lib.cv_last_error().decode())
        return inside.astype(bool)

    # -- text --
    def set_text_instancing(self, on):
        """Text draws upload glyph instances expanded on the device (default) or host-lowered outlines."""
        self.flush()
        self._lib.cv_set_text_instancing(self._h, 1 if on else 0)

    def set_font(self, font, size):
        if font is None or len(font) == 0:
            self._w.floats("SET_FONT_RESIZE", size)
            return None
        n_before = len(self.queries)
        self._w.floats("SET_FONT", size); self._w.raw("B", 1); self._w.blob(font)
        self.flush()
        return bool(self.queries[n_before][1]) if len(self.queries) > n_before else None

    def _text(self, op, text, x, y, maximum_width):
        self._w.floats(op, x, y, maximum_width)
        self._w.blob(text.encode("utf-8") if isinstance(text, str) else text)

    def fill_text(self, text, x, y, maximum_width=1.0e30): self._text("FILL_TEXT", text, x, y, maximum_width)
    def stroke_text(self, text, x, y, maximum_width=1.0e30): self._text("STROKE_TEXT", text, x, y, maximum_width)

    def measure_text(self, text):
        self.flush()
        return float(self._lib.cv_measure_text(self._h, text.encode("utf-8") if isinstance(text, str) else text))

    # -- images --
    def draw_image(self, image, width, height, stride, x, y, to_width, to_height):
        self._w.ints("DRAW_IMAGE", width, height, stride); self._w.raw("4f", x, y, to_width, to_height)
        self._w.blob(_image_bytes(image, width, height, stride))

    def get_image_data(self, width=None, height=None, x=0, y=0):
        """Returns an (height, width, 4) uint8 array (straight-alpha sRGB)."""
        width = self.width if width is None else width
        height = self.height if height is None else height
        self.flush()
        out = np.zeros((height, width, 4), np.uint8)
        rc = self._lib.cv_get_image_data(self._h, out.ctypes.data, width, height, 4 * width, x, y)
        if rc != 0:
            raise RuntimeError(self._lib.cv_last_error().decode())
        return out

    # -- beyond the reference API: the steps that follow get_image_data in its drivers, kept on the GPU --
    def device_canvas(self):
        """The cb200_canvas handle behind this canvas (None for tapped canvases)."""
        return self._lib.cv_device(self._h)

    def image_tensor(self, bgra=False):
        """get_image_data straight into a CUDA tensor: (rows, width, 4) uint8, no host copy."""
        import torch
        self.flush()
        rows = self.height if self._band is None else self._band[1]
        y0 = 0 if self._band is None else self._band[0]
        out = torch.empty((rows, self.width, 4), dtype=torch.uint8, device="cuda:%d" % self._device_index)
        dev = self.device_canvas()
        if self._lib.cb200_read_rgba8_into(dev, C.c_void_p(out.data_ptr()), self.width, rows, 0, y0) != 0 or \
                self._lib.cb200_sync(dev) != 0:
            raise RuntimeError(self._lib.cb200_last_error().decode())
        return out[..., [2, 1, 0, 3]] if bgra else out

    def framebuffer_tensor(self):
        """Zero-copy torch view of the linear premultiplied float framebuffer, (rows, width, 4) float32.
        Valid until the next draw call is flushed."""
        import torch
        self.flush()
        ptr, rows, width = C.c_void_p(), C.c_int(), C.c_int()
        if self._lib.cb200_framebuffer_device(self.device_canvas(), C.byref(ptr), C.byref(rows), C.byref(width)) != 0:
            raise RuntimeError(self._lib.cb200_last_error().decode())

        class _View:                                    # what torch.as_tensor needs to wrap foreign device memory
            __cuda_array_interface__ = {"shape": (rows.value, width.value, 4), "typestr": "<f4", "data": (ptr.value, False),
                                        "version": 2, "strides": None}
        return torch.as_tensor(_View(), device="cuda:%d" % self._device_index)

    def write_tga(self, path):
        """The image file demos/tiger/tiger.cpp:4333-4345 writes: 18-byte header + top-down BGRA rows,
        with the channel swap done by the readback kernel."""
        self.flush()
        if self._lib.cv_write_tga(self._h, str(path).encode()) != 0:
            raise RuntimeError(self._lib.cv_last_error().decode())

    def write_png(self, path):
        """The PNG the reference's test driver writes (write_png, test/test.cpp:2415-2507); pixels, framing and
        both checksums come off the device in one kernel (cb200_encode_png)."""
        self.flush()
        if self._lib.cv_write_png(self._h, str(path).encode()) != 0:
            raise RuntimeError(self._lib.cv_last_error().decode())

    def put_image_data(self, image, width, height, stride, x, y):
        self._w.ints("PUT_IMAGE_DATA", width, height, stride, x, y)
        self._w.blob(_image_bytes(image, width, height, stride))

    def read_f32(self, rows=None):
        """Linear premultiplied float framebuffer (rows, width, 4)."""
        self.flush()
        rows = self.height if rows is None else rows
        out = np.zeros((rows, self.width, 4), np.float32)
        rc = self._lib.cv_read_f32(self._h, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(self._lib.cv_last_error().decode())
        return out

    # -- state stack --
    def save(self):
        self._stack.append(dict(self._state)); self._w.bare("SAVE")

    def restore(self):
        if self._stack:
            self._state = self._stack.pop()
        self._w.bare("RESTORE")
