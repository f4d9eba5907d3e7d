// script.hpp -- "canvas script": a flat little-endian byte encoding of calls to
// the canvas_ity public API (reference src/canvas_ity.hpp:194-1148), one opcode
// per method or public data member.  It is how the Python mirror
// (canvas_ity_b200.Canvas) talks to the C++ front end in one call per flush, and
// how the reference's own drivers (test/test.cpp, demos/tiger/tiger.cpp) are
// captured once into tests/golden/*.cvs so the same call stream can be replayed
// into the reference build, the oracle and the B200 back end.
//
// run_script<Canvas>() is a template over any class with the canvas_ity API, so
// the identical replayer drives `canvas_ity::canvas` from either header.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace cb200_script {

enum op : uint8_t {
    OP_END = 0,
    OP_SCALE, OP_ROTATE, OP_TRANSLATE, OP_TRANSFORM, OP_SET_TRANSFORM,
    OP_SET_GLOBAL_ALPHA, OP_SET_COMPOSITE, OP_SET_SHADOW_COLOR, OP_SET_SHADOW_OFFSET_X,
    OP_SET_SHADOW_OFFSET_Y, OP_SET_SHADOW_BLUR, OP_SET_LINE_WIDTH, OP_SET_LINE_CAP,
    OP_SET_LINE_JOIN, OP_SET_MITER_LIMIT, OP_SET_LINE_DASH_OFFSET, OP_SET_LINE_DASH,
    OP_SET_COLOR, OP_SET_LINEAR_GRADIENT, OP_SET_RADIAL_GRADIENT, OP_ADD_COLOR_STOP,
    OP_SET_PATTERN, OP_BEGIN_PATH, OP_MOVE_TO, OP_CLOSE_PATH, OP_LINE_TO,
    OP_QUADRATIC_CURVE_TO, OP_BEZIER_CURVE_TO, OP_ARC_TO, OP_ARC, OP_RECTANGLE,
    OP_FILL, OP_STROKE, OP_CLIP, OP_IS_POINT_IN_PATH, OP_CLEAR_RECTANGLE,
    OP_FILL_RECTANGLE, OP_STROKE_RECTANGLE, OP_SET_TEXT_ALIGN, OP_SET_TEXT_BASELINE,
    OP_SET_FONT, OP_FILL_TEXT, OP_STROKE_TEXT, OP_MEASURE_TEXT, OP_DRAW_IMAGE,
    OP_GET_IMAGE_DATA, OP_PUT_IMAGE_DATA, OP_SAVE, OP_RESTORE,
    OP_SET_LINE_DASH_NULL, OP_SET_FONT_RESIZE,
    OP_COUNT
};

// ---- writer -------------------------------------------------------------------
struct writer {
    std::vector<uint8_t> bytes;
    void u8(uint8_t v) { bytes.push_back(v); }
    void u32(uint32_t v) { raw(&v, 4); }
    void i32(int32_t v) { raw(&v, 4); }
    void f32(float v) { raw(&v, 4); }
    void raw(const void *p, size_t n)
    {
        const uint8_t *b = static_cast<const uint8_t *>(p);
        bytes.insert(bytes.end(), b, b + n);
    }
    void blob(const void *p, size_t n) { u32(uint32_t(n)); raw(p, n); }
    void floats(uint8_t code, const float *v, int n)
    {
        u8(code);
        for (int i = 0; i < n; ++i) f32(v[i]);
    }
};

// ---- reader -------------------------------------------------------------------
struct reader {
    const uint8_t *p, *end;
    bool ok = true;
    bool need(size_t n) { if (size_t(end - p) < n) { ok = false; return false; } return true; }
    uint8_t u8() { if (!need(1)) return 0; return *p++; }
    uint32_t u32() { uint32_t v = 0; if (need(4)) { memcpy(&v, p, 4); p += 4; } return v; }
    int32_t i32() { return int32_t(u32()); }
    float f32() { float v = 0; if (need(4)) { memcpy(&v, p, 4); p += 4; } return v; }
    const uint8_t *blob(uint32_t &n)
    {
        n = u32();
        if (!need(n)) { n = 0; return p; }
        const uint8_t *b = p;
        p += n;
        return b;
    }
};

// Results of the synchronous queries met while replaying, in script order,
// with the value recorded when the script was captured (if any).
struct query_result {
    uint8_t code;         // OP_IS_POINT_IN_PATH / OP_MEASURE_TEXT / OP_GET_IMAGE_DATA / OP_SET_FONT
    float got, recorded;  // bool as 0/1; get_image_data: FNV-1a of the bytes, as float bits
    uint32_t got_bits, recorded_bits;
};

inline uint32_t fnv1a(const uint8_t *p, size_t n)
{
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 16777619u; }
    return h;
}

// Bytes an image argument of w x h pixels with row pitch `stride` spans, the way the reference indexes it
// (image[y * stride + 4 * x + c], hpp:2852-2860, 3313-3346, 3383-3408).  False when a script cannot carry it:
// a negative pitch has no meaning relative to the start of a blob.
inline bool image_span(int32_t w, int32_t h, int32_t stride, size_t &bytes)
{
    bytes = 0;
    if (w <= 0 || h <= 0) return true;                  // the call is a no-op in the reference; nothing is read
    if (stride < 0) return false;
    bytes = size_t(h - 1) * size_t(stride) + 4 * size_t(w);
    return true;
}

// Replays `bytes` into `cv`.  Returns the number of ops executed, or -1 on a
// malformed script.  get_image_data inside a script reads into a scratch buffer
// (tests that draw based on the pixels they read were resolved at capture time).
template <class Canvas, class Ns>
long run_script(Canvas &cv, const uint8_t *bytes, size_t size, std::vector<query_result> *queries)
{
    reader r = { bytes, bytes + size };
    long executed = 0;
    std::vector<uint8_t> scratch;
    while (r.ok && r.p < r.end) {
        uint8_t code = r.u8();
        if (code == OP_END) break;
        float a[8];
        // every op reads ALL its arguments first; a truncated or inconsistent record ends the replay with -1
        // before anything is dispatched
#define READ_OK do { if (!r.ok) return -1; } while (0)
        switch (code) {
        case OP_SCALE: a[0] = r.f32(); a[1] = r.f32(); READ_OK; cv.scale(a[0], a[1]); break;
        case OP_ROTATE: a[0] = r.f32(); READ_OK; cv.rotate(a[0]); break;
        case OP_TRANSLATE: a[0] = r.f32(); a[1] = r.f32(); READ_OK; cv.translate(a[0], a[1]); break;
        case OP_TRANSFORM:
            for (int i = 0; i < 6; ++i) a[i] = r.f32();
            READ_OK; cv.transform(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case OP_SET_TRANSFORM:
            for (int i = 0; i < 6; ++i) a[i] = r.f32();
            READ_OK; cv.set_transform(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case OP_SET_GLOBAL_ALPHA: a[0] = r.f32(); READ_OK; cv.set_global_alpha(a[0]); break;
        case OP_SET_COMPOSITE:
            { int32_t v = r.i32(); READ_OK; cv.global_composite_operation = static_cast<typename Ns::composite_operation>(v); } break;
        case OP_SET_SHADOW_COLOR:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.set_shadow_color(a[0], a[1], a[2], a[3]); break;
        case OP_SET_SHADOW_OFFSET_X: a[0] = r.f32(); READ_OK; cv.shadow_offset_x = a[0]; break;
        case OP_SET_SHADOW_OFFSET_Y: a[0] = r.f32(); READ_OK; cv.shadow_offset_y = a[0]; break;
        case OP_SET_SHADOW_BLUR: a[0] = r.f32(); READ_OK; cv.set_shadow_blur(a[0]); break;
        case OP_SET_LINE_WIDTH: a[0] = r.f32(); READ_OK; cv.set_line_width(a[0]); break;
        case OP_SET_LINE_CAP: { int32_t v = r.i32(); READ_OK; cv.line_cap = static_cast<typename Ns::cap_style>(v); } break;
        case OP_SET_LINE_JOIN: { int32_t v = r.i32(); READ_OK; cv.line_join = static_cast<typename Ns::join_style>(v); } break;
        case OP_SET_MITER_LIMIT: a[0] = r.f32(); READ_OK; cv.set_miter_limit(a[0]); break;
        case OP_SET_LINE_DASH_OFFSET: a[0] = r.f32(); READ_OK; cv.line_dash_offset = a[0]; break;
        case OP_SET_LINE_DASH: {
            int32_t n = r.i32();
            std::vector<float> seg;
            if (n > 0 && size_t(n) > size_t(r.end - r.p) / 4) return -1;       // more segments than bytes left
            for (int i = 0; i < n && r.ok; ++i) seg.push_back(r.f32());
            float dummy = 0.0f;
            READ_OK; cv.set_line_dash(seg.empty() ? &dummy : seg.data(), n);
            break;
        }
        case OP_SET_LINE_DASH_NULL: { int32_t n = r.i32(); READ_OK; cv.set_line_dash(0, n); break; }
        case OP_SET_COLOR: {
            int32_t which = r.i32();
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.set_color(static_cast<typename Ns::brush_type>(which), a[0], a[1], a[2], a[3]);
            break;
        }
        case OP_SET_LINEAR_GRADIENT: {
            int32_t which = r.i32();
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.set_linear_gradient(static_cast<typename Ns::brush_type>(which), a[0], a[1], a[2], a[3]);
            break;
        }
        case OP_SET_RADIAL_GRADIENT: {
            int32_t which = r.i32();
            for (int i = 0; i < 6; ++i) a[i] = r.f32();
            READ_OK; cv.set_radial_gradient(static_cast<typename Ns::brush_type>(which), a[0], a[1], a[2],
                                   a[3], a[4], a[5]);
            break;
        }
        case OP_ADD_COLOR_STOP: {
            int32_t which = r.i32();
            for (int i = 0; i < 5; ++i) a[i] = r.f32();
            READ_OK; cv.add_color_stop(static_cast<typename Ns::brush_type>(which), a[0], a[1], a[2], a[3], a[4]);
            break;
        }
        case OP_SET_PATTERN: {
            int32_t which = r.i32(), w = r.i32(), h = r.i32(), stride = r.i32(), rep = r.i32();
            uint32_t n;
            const uint8_t *img = r.blob(n);
            size_t span;
            if (!r.ok || !image_span(w, h, stride, span) || (n && n < span)) return -1;
            cv.set_pattern(static_cast<typename Ns::brush_type>(which), n ? img : 0, w, h, stride,
                           static_cast<typename Ns::repetition_style>(rep));
            break;
        }
        case OP_BEGIN_PATH: READ_OK; cv.begin_path(); break;
        case OP_MOVE_TO: a[0] = r.f32(); a[1] = r.f32(); READ_OK; cv.move_to(a[0], a[1]); break;
        case OP_CLOSE_PATH: READ_OK; cv.close_path(); break;
        case OP_LINE_TO: a[0] = r.f32(); a[1] = r.f32(); READ_OK; cv.line_to(a[0], a[1]); break;
        case OP_QUADRATIC_CURVE_TO:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.quadratic_curve_to(a[0], a[1], a[2], a[3]); break;
        case OP_BEZIER_CURVE_TO:
            for (int i = 0; i < 6; ++i) a[i] = r.f32();
            READ_OK; cv.bezier_curve_to(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case OP_ARC_TO:
            for (int i = 0; i < 5; ++i) a[i] = r.f32();
            READ_OK; cv.arc_to(a[0], a[1], a[2], a[3], a[4]); break;
        case OP_ARC: {
            for (int i = 0; i < 5; ++i) a[i] = r.f32();
            int32_t ccw = r.i32();
            READ_OK; cv.arc(a[0], a[1], a[2], a[3], a[4], ccw != 0);
            break;
        }
        case OP_RECTANGLE:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.rectangle(a[0], a[1], a[2], a[3]); break;
        case OP_FILL: READ_OK; cv.fill(); break;
        case OP_STROKE: READ_OK; cv.stroke(); break;
        case OP_CLIP: READ_OK; cv.clip(); break;
        case OP_IS_POINT_IN_PATH: {
            a[0] = r.f32(); a[1] = r.f32();
            uint8_t recorded = r.u8();
            READ_OK; bool got = cv.is_point_in_path(a[0], a[1]);
            if (queries) {
                query_result q = { code, got ? 1.0f : 0.0f, recorded ? 1.0f : 0.0f,
                                   got ? 1u : 0u, recorded ? 1u : 0u };
                queries->push_back(q);
            }
            break;
        }
        case OP_CLEAR_RECTANGLE:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.clear_rectangle(a[0], a[1], a[2], a[3]); break;
        case OP_FILL_RECTANGLE:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.fill_rectangle(a[0], a[1], a[2], a[3]); break;
        case OP_STROKE_RECTANGLE:
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            READ_OK; cv.stroke_rectangle(a[0], a[1], a[2], a[3]); break;
        case OP_SET_TEXT_ALIGN: { int32_t v = r.i32(); READ_OK; cv.text_align = static_cast<typename Ns::align_style>(v); } break;
        case OP_SET_TEXT_BASELINE:
            { int32_t v = r.i32(); READ_OK; cv.text_baseline = static_cast<typename Ns::baseline_style>(v); } break;
        case OP_SET_FONT: {
            float size_px = r.f32();
            uint8_t recorded = r.u8();
            uint32_t n;
            const uint8_t *font = r.blob(n);
            READ_OK; bool got = cv.set_font(n ? font : 0, int(n), size_px);
            if (queries) {
                query_result q = { code, got ? 1.0f : 0.0f, recorded ? 1.0f : 0.0f,
                                   got ? 1u : 0u, recorded ? 1u : 0u };
                queries->push_back(q);
            }
            break;
        }
        case OP_SET_FONT_RESIZE: {          // set_font(NULL, 0, size): keep the face
            float size_px = r.f32();
            READ_OK; cv.set_font(0, 0, size_px);
            break;
        }
        case OP_FILL_TEXT:
        case OP_STROKE_TEXT: {
            for (int i = 0; i < 3; ++i) a[i] = r.f32();
            uint32_t n;
            const uint8_t *s = r.blob(n);
            std::string text(reinterpret_cast<const char *>(s), n);
            READ_OK;
            if (code == OP_FILL_TEXT) cv.fill_text(text.c_str(), a[0], a[1], a[2]);
            else cv.stroke_text(text.c_str(), a[0], a[1], a[2]);
            break;
        }
        case OP_MEASURE_TEXT: {
            float recorded = r.f32();
            uint32_t n;
            const uint8_t *s = r.blob(n);
            std::string text(reinterpret_cast<const char *>(s), n);
            READ_OK; float got = cv.measure_text(text.c_str());
            if (queries) {
                query_result q = { code, got, recorded, 0, 0 };
                memcpy(&q.got_bits, &got, 4);
                memcpy(&q.recorded_bits, &recorded, 4);
                queries->push_back(q);
            }
            break;
        }
        case OP_DRAW_IMAGE: {
            int32_t w = r.i32(), h = r.i32(), stride = r.i32();
            for (int i = 0; i < 4; ++i) a[i] = r.f32();
            uint32_t n;
            const uint8_t *img = r.blob(n);
            size_t span;
            if (!r.ok || !image_span(w, h, stride, span) || (n && n < span)) return -1;
            cv.draw_image(n ? img : 0, w, h, stride, a[0], a[1], a[2], a[3]);
            break;
        }
        case OP_GET_IMAGE_DATA: {
            int32_t w = r.i32(), h = r.i32(), stride = r.i32(), x = r.i32(), y = r.i32();
            uint32_t recorded = r.u32();
            // the recorded hash covers h * stride bytes (capture_shim.hpp); the call writes up to image_span
            size_t span;
            if (!r.ok || !image_span(w, h, stride, span) || span > (size_t(1) << 33)) return -1;
            const size_t need = size_t(h > 0 ? h : 0) * size_t(stride > 0 ? stride : 0);
            scratch.assign((span > need ? span : need) + 16, 0);
            cv.get_image_data(scratch.data(), w, h, stride, x, y);
            if (queries) {
                uint32_t got = fnv1a(scratch.data(), need);
                query_result q = { code, 0.0f, 0.0f, got, recorded };
                queries->push_back(q);
            }
            break;
        }
        case OP_PUT_IMAGE_DATA: {
            int32_t w = r.i32(), h = r.i32(), stride = r.i32(), x = r.i32(), y = r.i32();
            uint32_t n;
            const uint8_t *img = r.blob(n);
            size_t span;
            if (!r.ok || !image_span(w, h, stride, span) || (n && n < span)) return -1;
            cv.put_image_data(n ? img : 0, w, h, stride, x, y);
            break;
        }
        case OP_SAVE: READ_OK; cv.save(); break;
        case OP_RESTORE: READ_OK; cv.restore(); break;
        default: return -1;
        }
#undef READ_OK
        if (!r.ok) return -1;
        ++executed;
    }
    return r.ok ? executed : -1;
}

}  // namespace cb200_script
