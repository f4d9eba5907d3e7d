// capture_shim.hpp -- TEST INFRASTRUCTURE, not product code.
//
// A `canvas_ity::canvas` look-alike that forwards every call to the real
// reference class (included beforehand under the name canvas_ity_ref) and at the
// same time records the call as a canvas script (script.hpp).  Compiling the
// reference's own drivers (test/test.cpp, demos/tiger/tiger.cpp -- read from
// /root/reference, never copied) against this shim yields, per test, the exact
// API call stream plus the reference's answers to the synchronous queries, which
// tools/make_golden.py commits as tests/golden/*.cvs fixtures.
#pragma once

#define CANVAS_ITY_B200_ENUMS_ONLY
#include "../include/canvas_ity.hpp"
#include "../canvas_ity_b200/csrc/host/script.hpp"

#include <vector>

namespace canvas_ity {

class canvas
{
public:
    composite_operation global_composite_operation;
    float shadow_offset_x, shadow_offset_y;
    cap_style line_cap;
    join_style line_join;
    float line_dash_offset;
    align_style text_align;
    baseline_style text_baseline;

    cb200_script::writer script;
    canvas_ity_ref::canvas real;
    int width_px, height_px;

    canvas(int w, int h) : real(w, h), width_px(w), height_px(h)
    {
        global_composite_operation = source_over;
        shadow_offset_x = shadow_offset_y = 0.0f;
        line_cap = butt; line_join = miter; line_dash_offset = 0.0f;
        text_align = start; text_baseline = alphabetic;
        seen_op = global_composite_operation; seen_sx = seen_sy = 0.0f;
        seen_cap = line_cap; seen_join = line_join; seen_dash = 0.0f;
        seen_align = text_align; seen_base = text_baseline;
    }
    ~canvas() { if (on_destroy) on_destroy(*this); }
    static void (*on_destroy)(canvas &);

    // public data members are sampled before every call so assignment order
    // relative to calls (save/restore included) is preserved in the script
    void sync()
    {
        using namespace cb200_script;
        if (global_composite_operation != seen_op) {
            seen_op = global_composite_operation;
            script.u8(OP_SET_COMPOSITE); script.i32(int(seen_op));
        }
        if (!same(shadow_offset_x, seen_sx)) { seen_sx = shadow_offset_x; script.u8(OP_SET_SHADOW_OFFSET_X); script.f32(seen_sx); }
        if (!same(shadow_offset_y, seen_sy)) { seen_sy = shadow_offset_y; script.u8(OP_SET_SHADOW_OFFSET_Y); script.f32(seen_sy); }
        if (line_cap != seen_cap) { seen_cap = line_cap; script.u8(OP_SET_LINE_CAP); script.i32(int(seen_cap)); }
        if (line_join != seen_join) { seen_join = line_join; script.u8(OP_SET_LINE_JOIN); script.i32(int(seen_join)); }
        if (!same(line_dash_offset, seen_dash)) { seen_dash = line_dash_offset; script.u8(OP_SET_LINE_DASH_OFFSET); script.f32(seen_dash); }
        if (text_align != seen_align) { seen_align = text_align; script.u8(OP_SET_TEXT_ALIGN); script.i32(int(seen_align)); }
        if (text_baseline != seen_base) { seen_base = text_baseline; script.u8(OP_SET_TEXT_BASELINE); script.i32(int(seen_base)); }
        real.global_composite_operation = canvas_ity_ref::composite_operation(int(global_composite_operation));
        real.shadow_offset_x = shadow_offset_x;
        real.shadow_offset_y = shadow_offset_y;
        real.line_cap = canvas_ity_ref::cap_style(int(line_cap));
        real.line_join = canvas_ity_ref::join_style(int(line_join));
        real.line_dash_offset = line_dash_offset;
        real.text_align = canvas_ity_ref::align_style(int(text_align));
        real.text_baseline = canvas_ity_ref::baseline_style(int(text_baseline));
    }
    void pull()      // restore() changes the members inside the real canvas
    {
        global_composite_operation = seen_op = composite_operation(int(real.global_composite_operation));
        shadow_offset_x = seen_sx = real.shadow_offset_x;
        shadow_offset_y = seen_sy = real.shadow_offset_y;
        line_cap = seen_cap = cap_style(int(real.line_cap));
        line_join = seen_join = join_style(int(real.line_join));
        line_dash_offset = seen_dash = real.line_dash_offset;
        text_align = seen_align = align_style(int(real.text_align));
        text_baseline = seen_base = baseline_style(int(real.text_baseline));
    }

#define CV_F(code, n, ...) { sync(); float v_[] = { __VA_ARGS__ }; script.floats(cb200_script::code, v_, n); }
    void scale(float x, float y) { CV_F(OP_SCALE, 2, x, y) real.scale(x, y); }
    void rotate(float a) { CV_F(OP_ROTATE, 1, a) real.rotate(a); }
    void translate(float x, float y) { CV_F(OP_TRANSLATE, 2, x, y) real.translate(x, y); }
    void transform(float a, float b, float c, float d, float e, float f) { CV_F(OP_TRANSFORM, 6, a, b, c, d, e, f) real.transform(a, b, c, d, e, f); }
    void set_transform(float a, float b, float c, float d, float e, float f) { CV_F(OP_SET_TRANSFORM, 6, a, b, c, d, e, f) real.set_transform(a, b, c, d, e, f); }
    void set_global_alpha(float a) { CV_F(OP_SET_GLOBAL_ALPHA, 1, a) real.set_global_alpha(a); }
    void set_shadow_color(float r, float g, float b, float a) { CV_F(OP_SET_SHADOW_COLOR, 4, r, g, b, a) real.set_shadow_color(r, g, b, a); }
    void set_shadow_blur(float l) { CV_F(OP_SET_SHADOW_BLUR, 1, l) real.set_shadow_blur(l); }
    void set_line_width(float w) { CV_F(OP_SET_LINE_WIDTH, 1, w) real.set_line_width(w); }
    void set_miter_limit(float l) { CV_F(OP_SET_MITER_LIMIT, 1, l) real.set_miter_limit(l); }
    void set_line_dash(float const *seg, int count)
    {
        sync();
        if (!seg) { script.u8(cb200_script::OP_SET_LINE_DASH_NULL); script.i32(count); }
        else {
            script.u8(cb200_script::OP_SET_LINE_DASH); script.i32(count);
            for (int i = 0; i < count; ++i) script.f32(seg[i]);
        }
        real.set_line_dash(seg, count);
    }
    void set_color(brush_type t, float r, float g, float b, float a)
    {
        sync(); script.u8(cb200_script::OP_SET_COLOR); script.i32(int(t));
        script.f32(r); script.f32(g); script.f32(b); script.f32(a);
        real.set_color(canvas_ity_ref::brush_type(int(t)), r, g, b, a);
    }
    void set_linear_gradient(brush_type t, float sx, float sy, float ex, float ey)
    {
        sync(); script.u8(cb200_script::OP_SET_LINEAR_GRADIENT); script.i32(int(t));
        script.f32(sx); script.f32(sy); script.f32(ex); script.f32(ey);
        real.set_linear_gradient(canvas_ity_ref::brush_type(int(t)), sx, sy, ex, ey);
    }
    void set_radial_gradient(brush_type t, float sx, float sy, float sr, float ex, float ey, float er)
    {
        sync(); script.u8(cb200_script::OP_SET_RADIAL_GRADIENT); script.i32(int(t));
        script.f32(sx); script.f32(sy); script.f32(sr); script.f32(ex); script.f32(ey); script.f32(er);
        real.set_radial_gradient(canvas_ity_ref::brush_type(int(t)), sx, sy, sr, ex, ey, er);
    }
    void add_color_stop(brush_type t, float o, float r, float g, float b, float a)
    {
        sync(); script.u8(cb200_script::OP_ADD_COLOR_STOP); script.i32(int(t));
        script.f32(o); script.f32(r); script.f32(g); script.f32(b); script.f32(a);
        real.add_color_stop(canvas_ity_ref::brush_type(int(t)), o, r, g, b, a);
    }
    static size_t image_bytes(unsigned char const *image, int w, int h, int stride)
    {
        if (!image) return 0;
        if (w <= 0 || h <= 0 || stride < 0) return 1;     // non-null but never read
        return size_t(h - 1) * size_t(stride) + size_t(w) * 4;
    }
    void set_pattern(brush_type t, unsigned char const *image, int w, int h, int stride, repetition_style rep)
    {
        sync(); script.u8(cb200_script::OP_SET_PATTERN); script.i32(int(t));
        script.i32(w); script.i32(h); script.i32(stride); script.i32(int(rep));
        script.blob(image, image_bytes(image, w, h, stride));
        real.set_pattern(canvas_ity_ref::brush_type(int(t)), image, w, h, stride,
                         canvas_ity_ref::repetition_style(int(rep)));
    }
    void begin_path() { sync(); script.u8(cb200_script::OP_BEGIN_PATH); real.begin_path(); }
    void move_to(float x, float y) { CV_F(OP_MOVE_TO, 2, x, y) real.move_to(x, y); }
    void close_path() { sync(); script.u8(cb200_script::OP_CLOSE_PATH); real.close_path(); }
    void line_to(float x, float y) { CV_F(OP_LINE_TO, 2, x, y) real.line_to(x, y); }
    void quadratic_curve_to(float cx, float cy, float x, float y) { CV_F(OP_QUADRATIC_CURVE_TO, 4, cx, cy, x, y) real.quadratic_curve_to(cx, cy, x, y); }
    void bezier_curve_to(float a, float b, float c, float d, float x, float y) { CV_F(OP_BEZIER_CURVE_TO, 6, a, b, c, d, x, y) real.bezier_curve_to(a, b, c, d, x, y); }
    void arc_to(float vx, float vy, float x, float y, float r) { CV_F(OP_ARC_TO, 5, vx, vy, x, y, r) real.arc_to(vx, vy, x, y, r); }
    void arc(float x, float y, float r, float a0, float a1, bool ccw = false)
    {
        CV_F(OP_ARC, 5, x, y, r, a0, a1) script.i32(ccw ? 1 : 0);
        real.arc(x, y, r, a0, a1, ccw);
    }
    void rectangle(float x, float y, float w, float h) { CV_F(OP_RECTANGLE, 4, x, y, w, h) real.rectangle(x, y, w, h); }
    void fill() { sync(); script.u8(cb200_script::OP_FILL); real.fill(); }
    void stroke() { sync(); script.u8(cb200_script::OP_STROKE); real.stroke(); }
    void clip() { sync(); script.u8(cb200_script::OP_CLIP); real.clip(); }
    bool is_point_in_path(float x, float y)
    {
        sync();
        bool hit = real.is_point_in_path(x, y);
        script.u8(cb200_script::OP_IS_POINT_IN_PATH); script.f32(x); script.f32(y); script.u8(hit ? 1 : 0);
        return hit;
    }
    void clear_rectangle(float x, float y, float w, float h) { CV_F(OP_CLEAR_RECTANGLE, 4, x, y, w, h) real.clear_rectangle(x, y, w, h); }
    void fill_rectangle(float x, float y, float w, float h) { CV_F(OP_FILL_RECTANGLE, 4, x, y, w, h) real.fill_rectangle(x, y, w, h); }
    void stroke_rectangle(float x, float y, float w, float h) { CV_F(OP_STROKE_RECTANGLE, 4, x, y, w, h) real.stroke_rectangle(x, y, w, h); }
    bool set_font(unsigned char const *font, int bytes, float size)
    {
        sync();
        bool ok = real.set_font(font, bytes, size);
        if (font && bytes) {
            script.u8(cb200_script::OP_SET_FONT); script.f32(size); script.u8(ok ? 1 : 0);
            script.blob(font, size_t(bytes > 0 ? bytes : 0));
        } else {
            script.u8(cb200_script::OP_SET_FONT_RESIZE); script.f32(size);
        }
        return ok;
    }
    void text_call(uint8_t code, char const *text, float x, float y, float mw)
    {
        script.u8(code); script.f32(x); script.f32(y); script.f32(mw);
        size_t n = 0;
        while (text && text[n]) ++n;
        script.blob(text, n);
    }
    void fill_text(char const *text, float x, float y, float mw = 1.0e30f)
    {
        sync(); text_call(cb200_script::OP_FILL_TEXT, text, x, y, mw);
        real.fill_text(text, x, y, mw);
    }
    void stroke_text(char const *text, float x, float y, float mw = 1.0e30f)
    {
        sync(); text_call(cb200_script::OP_STROKE_TEXT, text, x, y, mw);
        real.stroke_text(text, x, y, mw);
    }
    float measure_text(char const *text)
    {
        sync();
        float w = real.measure_text(text);
        script.u8(cb200_script::OP_MEASURE_TEXT); script.f32(w);
        size_t n = 0;
        while (text && text[n]) ++n;
        script.blob(text, n);
        return w;
    }
    void draw_image(unsigned char const *image, int w, int h, int stride, float x, float y, float tw, float th)
    {
        sync(); script.u8(cb200_script::OP_DRAW_IMAGE);
        script.i32(w); script.i32(h); script.i32(stride);
        script.f32(x); script.f32(y); script.f32(tw); script.f32(th);
        script.blob(image, image_bytes(image, w, h, stride));
        real.draw_image(image, w, h, stride, x, y, tw, th);
    }
    void get_image_data(unsigned char *image, int w, int h, int stride, int x, int y)
    {
        sync();
        real.get_image_data(image, w, h, stride, x, y);
        if (!image) return;                              // documented no-op
#ifdef CAPTURE_SKIP_GET_IMAGE_DATA
        return;                                          // driver's own final readback
#endif
        // record what a zero-initialised destination of h*stride bytes looks like
        size_t need = size_t(h > 0 ? h : 0) * size_t(stride > 0 ? stride : 0);
        std::vector<unsigned char> scratch(need + 16, 0);
        real.get_image_data(scratch.data(), w, h, stride, x, y);
        script.u8(cb200_script::OP_GET_IMAGE_DATA);
        script.i32(w); script.i32(h); script.i32(stride); script.i32(x); script.i32(y);
        script.u32(cb200_script::fnv1a(scratch.data(), need));
    }
    void put_image_data(unsigned char const *image, int w, int h, int stride, int x, int y)
    {
        sync(); script.u8(cb200_script::OP_PUT_IMAGE_DATA);
        script.i32(w); script.i32(h); script.i32(stride); script.i32(x); script.i32(y);
        script.blob(image, image_bytes(image, w, h, stride));
        real.put_image_data(image, w, h, stride, x, y);
    }
    void save() { sync(); script.u8(cb200_script::OP_SAVE); real.save(); }
    void restore() { sync(); script.u8(cb200_script::OP_RESTORE); real.restore(); pull(); }
#undef CV_F

private:
    static bool same(float a, float b) { return a == b || (a != a && b != b); }
    composite_operation seen_op; float seen_sx, seen_sy; cap_style seen_cap; join_style seen_join;
    float seen_dash; align_style seen_align; baseline_style seen_base;
    canvas(canvas const &);
    canvas &operator=(canvas const &);
};

void (*canvas::on_destroy)(canvas &) = 0;

}  // namespace canvas_ity
