"""Canvas scripts: the byte encoding of canvas_ity API calls (csrc/host/script.hpp)."""
import struct

OPS = ["END", "SCALE", "ROTATE", "TRANSLATE", "TRANSFORM", "SET_TRANSFORM", "SET_GLOBAL_ALPHA", "SET_COMPOSITE",
       "SET_SHADOW_COLOR", "SET_SHADOW_OFFSET_X", "SET_SHADOW_OFFSET_Y", "SET_SHADOW_BLUR", "SET_LINE_WIDTH",
       "SET_LINE_CAP", "SET_LINE_JOIN", "SET_MITER_LIMIT", "SET_LINE_DASH_OFFSET", "SET_LINE_DASH", "SET_COLOR",
       "SET_LINEAR_GRADIENT", "SET_RADIAL_GRADIENT", "ADD_COLOR_STOP", "SET_PATTERN", "BEGIN_PATH", "MOVE_TO",
       "CLOSE_PATH", "LINE_TO", "QUADRATIC_CURVE_TO", "BEZIER_CURVE_TO", "ARC_TO", "ARC", "RECTANGLE", "FILL",
       "STROKE", "CLIP", "IS_POINT_IN_PATH", "CLEAR_RECTANGLE", "FILL_RECTANGLE", "STROKE_RECTANGLE",
       "SET_TEXT_ALIGN", "SET_TEXT_BASELINE", "SET_FONT", "FILL_TEXT", "STROKE_TEXT", "MEASURE_TEXT", "DRAW_IMAGE",
       "GET_IMAGE_DATA", "PUT_IMAGE_DATA", "SAVE", "RESTORE", "SET_LINE_DASH_NULL", "SET_FONT_RESIZE"]
OP = {name: i for i, name in enumerate(OPS)}


class ScriptWriter:
    """Append-only encoder; one method per opcode family."""

    def __init__(self):
        self.buf = bytearray()

    def floats(self, op, *vals):
        self.buf += struct.pack("<B%df" % len(vals), OP[op], *vals)

    def ints(self, op, *vals):
        self.buf += struct.pack("<B%di" % len(vals), OP[op], *vals)

    def bare(self, op):
        self.buf.append(OP[op])

    def raw(self, fmt, *vals):
        self.buf += struct.pack("<" + fmt, *vals)

    def blob(self, data):
        data = bytes(data) if data is not None else b""
        self.buf += struct.pack("<I", len(data)) + data

    def take(self):
        out = bytes(self.buf)
        self.buf = bytearray()
        return out
