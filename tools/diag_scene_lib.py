"""Canvas-script splitting helpers shared by the diag tools."""
import struct
from canvas_ity_b200.script import OPS

FMT = {'SCALE': 8, 'ROTATE': 4, 'TRANSLATE': 8, 'SET_GLOBAL_ALPHA': 4, 'SET_COMPOSITE': 4, 'SET_SHADOW_COLOR': 16, 'SET_SHADOW_OFFSET_X': 4,
       'SET_SHADOW_OFFSET_Y': 4, 'SET_SHADOW_BLUR': 4, 'SET_LINE_WIDTH': 4, 'SET_LINE_CAP': 4, 'SET_LINE_JOIN': 4, 'SET_MITER_LIMIT': 4,
       'SET_LINE_DASH_OFFSET': 4, 'SET_COLOR': 20, 'SET_LINEAR_GRADIENT': 20, 'SET_RADIAL_GRADIENT': 28, 'ADD_COLOR_STOP': 24, 'BEGIN_PATH': 0,
       'MOVE_TO': 8, 'CLOSE_PATH': 0, 'LINE_TO': 8, 'QUADRATIC_CURVE_TO': 16, 'BEZIER_CURVE_TO': 24, 'ARC': 24, 'ARC_TO': 20, 'RECTANGLE': 16,
       'FILL': 0, 'STROKE': 0, 'CLIP': 0, 'FILL_RECTANGLE': 16, 'STROKE_RECTANGLE': 16, 'CLEAR_RECTANGLE': 16, 'SAVE': 0, 'RESTORE': 0,
       'SET_TEXT_ALIGN': 4, 'SET_TEXT_BASELINE': 4}
INTS = ('SET_COMPOSITE', 'SET_LINE_CAP', 'SET_LINE_JOIN', 'SET_TEXT_ALIGN', 'SET_TEXT_BASELINE')
DRAWS = ('FILL', 'STROKE', 'FILL_RECTANGLE', 'STROKE_RECTANGLE', 'CLEAR_RECTANGLE', 'FILL_TEXT', 'STROKE_TEXT', 'CLIP', 'DRAW_IMAGE', 'PUT_IMAGE_DATA')


def split(s):
    ops, i = [], 0
    while i < len(s):
        name, j = OPS[s[i]], i + 1
        if name in FMT: j += FMT[name]
        elif name == 'SET_LINE_DASH': j += 4 + 4 * struct.unpack_from('<i', s, j)[0]
        elif name == 'SET_PATTERN': j += 20; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'SET_FONT': j += 5; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name in ('FILL_TEXT', 'STROKE_TEXT'): j += 12; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'DRAW_IMAGE': j += 28; j += 4 + struct.unpack_from('<I', s, j)[0]
        elif name == 'PUT_IMAGE_DATA': j += 20; j += 4 + struct.unpack_from('<I', s, j)[0]
        else: raise SystemExit('unknown op ' + name)
        ops.append((name, i, j)); i = j
    return ops


def show(s, name, b, e):
    if name in INTS: return struct.unpack_from('<i', s, b + 1)
    if name in FMT and FMT[name]:
        v = struct.unpack_from('<%df' % (FMT[name] // 4), s, b + 1)
        return tuple(round(x, 2) for x in v)
    if name == 'SET_LINE_DASH': return struct.unpack_from('<i', s, b + 1)
    return ''


