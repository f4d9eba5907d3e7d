"""HTML5 ``CanvasRenderingContext2D``-style facade over :class:`canvas_ity_b200.Canvas`.

The reference models its C++ API on the W3C 2D canvas specification (README.md:23-25: the differences are
"mostly syntactic", so that a thin JavaScript binding can be put on top).  This is that thin binding for
Python callers: camelCase methods, string-valued properties (``fillStyle = "#3a7"``, ``lineCap = "round"``,
``globalCompositeOperation = "destination-out"``, ``font = "24px sans"``), gradient / pattern objects,
``measureText(...).width``.  Every call maps one-to-one onto the canvas_ity method of the same meaning
(reference src/canvas_ity.hpp, the doc comment of each method names its HTML5 counterpart), so nothing here
renders -- it records into the same canvas script as ``Canvas``.

    ctx = Context2D(256, 256, fonts={"sans": open("font.ttf", "rb").read()})
    ctx.fillStyle = "rgba(30, 120, 200, 0.8)"
    ctx.beginPath(); ctx.arc(128, 128, 90, 0, 6.2832); ctx.fill()
    ctx.font = "32px sans"; ctx.fillStyle = "black"; ctx.fillText("hello", 60, 140)
    rgba = ctx.getImageData(0, 0, 256, 256)
"""
import re

import numpy as np

from . import canvas as cv
from .canvas import Canvas

_OPS = {"source-in": cv.source_in, "copy": cv.source_copy, "source-out": cv.source_out,
        "destination-in": cv.destination_in, "destination-atop": cv.destination_atop, "lighter": cv.lighter,
        "destination-over": cv.destination_over, "destination-out": cv.destination_out,
        "source-atop": cv.source_atop, "source-over": cv.source_over, "xor": cv.exclusive_or}
_CAPS = {"butt": cv.butt, "square": cv.square, "round": cv.circle}
_JOINS = {"miter": cv.miter, "bevel": cv.bevel, "round": cv.rounded}
_ALIGN = {"left": cv.leftward, "right": cv.rightward, "center": cv.center, "start": cv.start, "end": cv.ending}
_BASELINE = {"alphabetic": cv.alphabetic, "top": cv.top, "middle": cv.middle, "bottom": cv.bottom,
             "hanging": cv.hanging, "ideographic": cv.ideographic}
_REPEAT = {"repeat": cv.repeat, "repeat-x": cv.repeat_x, "repeat-y": cv.repeat_y, "no-repeat": cv.no_repeat,
           "": cv.repeat, None: cv.repeat}
_NAMED = {"black": (0, 0, 0), "white": (255, 255, 255), "red": (255, 0, 0), "lime": (0, 255, 0), "green": (0, 128, 0),
          "blue": (0, 0, 255), "yellow": (255, 255, 0), "cyan": (0, 255, 255), "aqua": (0, 255, 255),
          "magenta": (255, 0, 255), "fuchsia": (255, 0, 255), "gray": (128, 128, 128), "grey": (128, 128, 128),
          "silver": (192, 192, 192), "maroon": (128, 0, 0), "olive": (128, 128, 0), "navy": (0, 0, 128),
          "purple": (128, 0, 128), "teal": (0, 128, 128), "orange": (255, 165, 0)}


def parse_color(text):
    """CSS colour -> (r, g, b, a) sRGB floats in [0, 1] (what set_color takes, hpp:478), or None if unparsable
    (HTML5: an invalid value leaves the property unchanged)."""
    if isinstance(text, (tuple, list)) and len(text) in (3, 4):
        return tuple(float(v) for v in text) + ((1.0,) if len(text) == 3 else ())
    if not isinstance(text, str):
        return None
    t = text.strip().lower()
    if t == "transparent":
        return (0.0, 0.0, 0.0, 0.0)
    if t in _NAMED:
        r, g, b = _NAMED[t]
        return (r / 255.0, g / 255.0, b / 255.0, 1.0)
    m = re.fullmatch(r"#([0-9a-f]{3,8})", t)
    if m and len(m.group(1)) in (3, 4, 6, 8):
        h = m.group(1)
        if len(h) <= 4:
            h = "".join(c * 2 for c in h)
        vals = [int(h[i:i + 2], 16) / 255.0 for i in range(0, len(h), 2)]
        return tuple(vals) + ((1.0,) if len(vals) == 3 else ())
    m = re.fullmatch(r"rgba?\(([^)]*)\)", t)
    if m:
        parts = [p for p in re.split(r"[,\s/]+", m.group(1).strip()) if p]
        if len(parts) in (3, 4):
            try:
                rgb = [float(p[:-1]) / 100.0 if p.endswith("%") else float(p) / 255.0 for p in parts[:3]]
                a = 1.0 if len(parts) == 3 else (float(parts[3][:-1]) / 100.0 if parts[3].endswith("%") else float(parts[3]))
            except ValueError:
                return None
            clamp = lambda v: min(max(v, 0.0), 1.0)
            return (clamp(rgb[0]), clamp(rgb[1]), clamp(rgb[2]), clamp(a))
    return None


class CanvasGradient:
    def __init__(self, kind, args):
        self.kind, self.args, self.stops = kind, tuple(float(a) for a in args), []

    def addColorStop(self, offset, color):
        c = parse_color(color)
        if c is not None:
            self.stops.append((float(offset), c))


class CanvasPattern:
    def __init__(self, image, repetition):
        img = np.ascontiguousarray(np.asarray(image, np.uint8))
        if img.ndim != 3 or img.shape[2] != 4:
            raise ValueError("createPattern: image must be (height, width, 4) uint8 RGBA")
        self.image, self.repetition = img, _REPEAT[repetition]


class TextMetrics:
    def __init__(self, width):
        self.width = width


class Context2D:
    def __init__(self, width, height, fonts=None, canvas=None, **canvas_args):
        object.__setattr__(self, "_c", canvas if canvas is not None else Canvas(width, height, **canvas_args))
        object.__setattr__(self, "_fonts", dict(fonts or {}))
        object.__setattr__(self, "_p", {
            "fillStyle": "#000000", "strokeStyle": "#000000", "lineWidth": 1.0, "lineCap": "butt", "lineJoin": "miter",
            "miterLimit": 10.0, "lineDashOffset": 0.0, "globalAlpha": 1.0, "globalCompositeOperation": "source-over",
            "shadowColor": "rgba(0, 0, 0, 0)", "shadowBlur": 0.0, "shadowOffsetX": 0.0, "shadowOffsetY": 0.0,
            "font": "10px sans-serif", "textAlign": "start", "textBaseline": "alphabetic"})
        object.__setattr__(self, "_dash", [])
        object.__setattr__(self, "_stack", [])
        object.__setattr__(self, "_face", None)

    canvas = property(lambda self: self._c)

    # ---- properties ----
    def __getattr__(self, name):
        props = object.__getattribute__(self, "_p")
        if name in props:
            return props[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in self._p:
            raise AttributeError("Context2D has no property %r" % name)
        if self._apply(name, value):
            self._p[name] = value

    def _style(self, which, value):
        c = self._c
        if isinstance(value, CanvasGradient):
            if value.kind == "linear":
                c.set_linear_gradient(which, *value.args)
            else:
                c.set_radial_gradient(which, *value.args)
            for offset, col in value.stops:
                c.add_color_stop(which, offset, *col)
            return True
        if isinstance(value, CanvasPattern):
            h, w, _ = value.image.shape
            c.set_pattern(which, value.image, w, h, 4 * w, value.repetition)
            return True
        col = parse_color(value)
        if col is None:
            return False
        c.set_color(which, *col)
        return True

    def _apply(self, name, value):
        c = self._c
        if name == "fillStyle":
            return self._style(cv.fill_style, value)
        if name == "strokeStyle":
            return self._style(cv.stroke_style, value)
        if name == "lineWidth":
            if not value > 0:
                return False
            c.set_line_width(float(value))
        elif name == "lineCap":
            if value not in _CAPS:
                return False
            c.line_cap = _CAPS[value]
        elif name == "lineJoin":
            if value not in _JOINS:
                return False
            c.line_join = _JOINS[value]
        elif name == "miterLimit":
            if not value > 0:
                return False
            c.set_miter_limit(float(value))
        elif name == "lineDashOffset":
            c.line_dash_offset = float(value)
        elif name == "globalAlpha":
            if not 0.0 <= value <= 1.0:
                return False
            c.set_global_alpha(float(value))
        elif name == "globalCompositeOperation":
            if value not in _OPS:
                return False
            c.global_composite_operation = _OPS[value]
        elif name == "shadowColor":
            col = parse_color(value)
            if col is None:
                return False
            c.set_shadow_color(*col)
        elif name == "shadowBlur":
            if not value >= 0:
                return False
            c.set_shadow_blur(float(value))
        elif name == "shadowOffsetX":
            c.shadow_offset_x = float(value)
        elif name == "shadowOffsetY":
            c.shadow_offset_y = float(value)
        elif name == "font":
            m = re.search(r"(\d+(?:\.\d+)?)px\s+(.*)$", str(value).strip())
            if not m:
                return False
            size, family = float(m.group(1)), m.group(2).strip().strip("'\"")
            data = self._fonts.get(family)
            if data is None and self._face is None:
                return False                                   # no such face registered (fonts={...})
            if data is not None and family != self._face:
                if c.set_font(data, size) is False:
                    return False
                object.__setattr__(self, "_face", family)
            else:
                c.set_font(None, size)                         # same face, new size
        elif name == "textAlign":
            if value not in _ALIGN:
                return False
            c.text_align = _ALIGN[value]
        elif name == "textBaseline":
            if value not in _BASELINE:
                return False
            c.text_baseline = _BASELINE[value]
        return True

    # ---- state ----
    def save(self):
        self._stack.append((dict(self._p), list(self._dash), self._face))
        self._c.save()

    def restore(self):
        if not self._stack:
            return
        props, dash, face = self._stack.pop()
        self._p.update(props)
        object.__setattr__(self, "_dash", dash)
        object.__setattr__(self, "_face", face)
        self._c.restore()

    # ---- transforms ----
    def scale(self, x, y): self._c.scale(x, y)
    def rotate(self, angle): self._c.rotate(angle)
    def translate(self, x, y): self._c.translate(x, y)
    def transform(self, a, b, c, d, e, f): self._c.transform(a, b, c, d, e, f)
    def setTransform(self, a, b, c, d, e, f): self._c.set_transform(a, b, c, d, e, f)
    def resetTransform(self): self._c.set_transform(1, 0, 0, 1, 0, 0)

    # ---- styles ----
    def createLinearGradient(self, x0, y0, x1, y1): return CanvasGradient("linear", (x0, y0, x1, y1))
    def createRadialGradient(self, x0, y0, r0, x1, y1, r1): return CanvasGradient("radial", (x0, y0, r0, x1, y1, r1))
    def createPattern(self, image, repetition="repeat"): return CanvasPattern(image, repetition)

    def setLineDash(self, segments):
        segs = [float(s) for s in segments]
        if any(s < 0 or s != s for s in segs):
            return
        object.__setattr__(self, "_dash", segs * (2 if len(segs) % 2 else 1))
        self._c.set_line_dash(segs)

    def getLineDash(self): return list(self._dash)

    # ---- paths ----
    def beginPath(self): self._c.begin_path()
    def closePath(self): self._c.close_path()
    def moveTo(self, x, y): self._c.move_to(x, y)
    def lineTo(self, x, y): self._c.line_to(x, y)
    def quadraticCurveTo(self, cpx, cpy, x, y): self._c.quadratic_curve_to(cpx, cpy, x, y)
    def bezierCurveTo(self, c1x, c1y, c2x, c2y, x, y): self._c.bezier_curve_to(c1x, c1y, c2x, c2y, x, y)
    def arcTo(self, x1, y1, x2, y2, radius): self._c.arc_to(x1, y1, x2, y2, radius)
    def arc(self, x, y, radius, start, end, counterclockwise=False): self._c.arc(x, y, radius, start, end, counterclockwise)
    def rect(self, x, y, w, h): self._c.rectangle(x, y, w, h)
    def fill(self): self._c.fill()
    def stroke(self): self._c.stroke()
    def clip(self): self._c.clip()
    def isPointInPath(self, x, y): return self._c.is_point_in_path(x, y)
    def arePointsInPath(self, xy): return self._c.points_in_path(xy)       # extension: one device launch for many points

    # ---- rectangles, text, images ----
    def clearRect(self, x, y, w, h): self._c.clear_rectangle(x, y, w, h)
    def fillRect(self, x, y, w, h): self._c.fill_rectangle(x, y, w, h)
    def strokeRect(self, x, y, w, h): self._c.stroke_rectangle(x, y, w, h)

    def fillText(self, text, x, y, maxWidth=None): self._c.fill_text(text, x, y, 1.0e30 if maxWidth is None else maxWidth)
    def strokeText(self, text, x, y, maxWidth=None): self._c.stroke_text(text, x, y, 1.0e30 if maxWidth is None else maxWidth)
    def measureText(self, text): return TextMetrics(self._c.measure_text(text))

    def drawImage(self, image, dx, dy, dw=None, dh=None):
        img = np.ascontiguousarray(np.asarray(image, np.uint8))
        h, w, _ = img.shape
        self._c.draw_image(img, w, h, 4 * w, dx, dy, w if dw is None else dw, h if dh is None else dh)

    def getImageData(self, sx, sy, sw, sh): return self._c.get_image_data(int(sw), int(sh), int(sx), int(sy))

    def putImageData(self, image, dx, dy):
        img = np.ascontiguousarray(np.asarray(image, np.uint8))
        h, w, _ = img.shape
        self._c.put_image_data(img, w, h, 4 * w, int(dx), int(dy))

    def close(self): self._c.close()
