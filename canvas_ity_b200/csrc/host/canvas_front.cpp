// canvas_front.cpp -- host side of the drop-in canvas_ity::canvas.
//
// What runs here is exactly what BASELINE.json's north_star leaves on the CPU:
// state setters (reference src/canvas_ity.hpp "hpp":2657-2864), path building
// (hpp:2866-3042), TTF parsing / text layout (hpp:1533-1846, 3205-3311) and the
// save stack (hpp:3410-3465).  Draw entry points (hpp:3044-3203, 3276-3346) do
// not rasterise: they lower the call to a cb200_draw + pooled geometry and queue
// it.  Host math keeps the reference's operation order so device-space control
// points are bit-identical to the reference's `path` contents.
#include "front_state.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <stdexcept>
#include <unordered_map>

namespace canvas_ity {

using cb200::apply;
using cb200::dot;
using cb200::mix;
using cb200::perp;
using cb200::unit;
using cb200::v2;

// ---------------------------------------------------------------- colour ----

static float srgb_decode(float v)       // hpp:1273-1275
{
    return v < 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
}

static color4 linear_of(float r, float g, float b, float a)
{
    color4 c = { srgb_decode(cb200::clamp01(r)), srgb_decode(cb200::clamp01(g)),
                 srgb_decode(cb200::clamp01(b)), cb200::clamp01(a) };
    return c;
}

color4 srgb_to_premultiplied_linear(float r, float g, float b, float a)
{
    color4 c = linear_of(r, g, b, a);
    c.r *= c.a; c.g *= c.a; c.b *= c.a;
    return c;
}

static color4 texel_to_premultiplied_linear(const uint8_t *t)   // hpp:2856-2859
{
    float a = t[3] / 255.0f;
    color4 c = { srgb_decode(t[0] / 255.0f) * a, srgb_decode(t[1] / 255.0f) * a,
                 srgb_decode(t[2] / 255.0f) * a, a };
    return c;
}

// ------------------------------------------------------------- lifecycle ----

static const affine k_identity = { 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f };

static void init_state(canvas &self, canvas::host_state *s, int width, int height)
{
    s->width = width;
    s->height = height;
    s->forward = k_identity;
    s->inverse = k_identity;
    self.global_composite_operation = source_over;
    self.shadow_offset_x = 0.0f;
    self.shadow_offset_y = 0.0f;
    self.line_cap = butt;
    self.line_join = miter;
    self.line_dash_offset = 0.0f;
    self.text_align = start;
    self.text_baseline = alphabetic;
    self.set_color(fill_style, 0.0f, 0.0f, 0.0f, 1.0f);
    self.set_color(stroke_style, 0.0f, 0.0f, 0.0f, 1.0f);
}

canvas::canvas(int width, int height) : self(new host_state)
{
    init_state(*this, self, width, height);
    int rc = cb200_canvas_create(width, height, 0, &self->device);
    if (rc != CB200_OK) {
        std::string why = std::string("canvas_b200: cannot create device canvas: ") +
                          cb200_last_error();
        delete self;
        throw std::runtime_error(why);     // there is no CPU fallback
    }
}

canvas::canvas(int width, int height, int device, int band_y0, int band_rows)
    : self(new host_state)
{
    init_state(*this, self, width, height);
    if (device < 0)
        return;                             // tap-only canvas: caller installs self->tap
    int rc = cb200_canvas_create_band(width, height, band_y0, band_rows, device,
                                      &self->device);
    if (rc != CB200_OK) {
        std::string why = std::string("canvas_b200: cannot create device canvas: ") +
                          cb200_last_error();
        delete self;
        throw std::runtime_error(why);
    }
}

canvas::~canvas()
{
    if (self->device) {
        self->flush();
        cb200_canvas_destroy(self->device);
    }
    delete self;
}

// ------------------------------------------------------------ transforms ----

void canvas::scale(float x, float y) { transform(x, 0.0f, 0.0f, y, 0.0f, 0.0f); }

void canvas::rotate(float angle)
{
    float c = cosf(angle), s = sinf(angle);
    transform(c, s, -s, c, 0.0f, 0.0f);
}

void canvas::translate(float x, float y) { transform(1.0f, 0.0f, 0.0f, 1.0f, x, y); }

void canvas::transform(float a, float b, float c, float d, float e, float f)
{
    const affine &m = self->forward;       // current * new (hpp:2687-2692)
    set_transform(m.a * a + m.c * b, m.b * a + m.d * b,
                  m.a * c + m.c * d, m.b * c + m.d * d,
                  m.a * e + m.c * f + m.e, m.b * e + m.d * f + m.f);
}

void canvas::set_transform(float a, float b, float c, float d, float e, float f)
{
    float det = a * d - b * c;
    float k = det != 0.0f ? 1.0f / det : 0.0f;      // singular -> zero inverse
    affine fwd = { a, b, c, d, e, f };
    affine inv = { k * d, k * -b, k * -c, k * a, k * (c * f - d * e), k * (b * e - a * f) };
    self->forward = fwd;
    self->inverse = inv;
}

static bool singular(const affine &m) { return m.a * m.d - m.b * m.c == 0.0f; }

// --------------------------------------------------------- simple setters ----

void canvas::set_global_alpha(float alpha)
{
    if (0.0f <= alpha && alpha <= 1.0f) self->global_alpha = alpha;
}

void canvas::set_shadow_color(float r, float g, float b, float a)
{
    self->shadow_color = srgb_to_premultiplied_linear(r, g, b, a);
}

void canvas::set_shadow_blur(float level) { if (0.0f <= level) self->shadow_blur = level; }
void canvas::set_line_width(float width) { if (0.0f < width) self->line_width = width; }
void canvas::set_miter_limit(float limit) { if (0.0f < limit) self->miter_limit = limit; }

void canvas::set_line_dash(float const *segments, int count)
{
    if (segments)
        for (int i = 0; i < count; ++i)
            if (segments[i] < 0.0f) return;          // any negative: ignore the call
    self->dash.clear();
    if (!segments) return;
    int copies = (count & 1) ? 2 : 1;                // odd lists are doubled
    for (int k = 0; k < copies; ++k)
        self->dash.insert(self->dash.end(), segments, segments + count);
}

static brush_state &pick(canvas::host_state *s, brush_type type)
{
    brush_state &b = type == fill_style ? s->fill : s->stroke;
    b.serial = ++s->serial_counter;
    return b;
}

void canvas::set_color(brush_type type, float r, float g, float b, float a)
{
    brush_state &br = pick(self, type);
    br.type = CB200_BRUSH_COLOR;
    br.colors.assign(1, srgb_to_premultiplied_linear(r, g, b, a));
}

void canvas::set_linear_gradient(brush_type type, float sx, float sy, float ex, float ey)
{
    brush_state &br = pick(self, type);
    br.type = CB200_BRUSH_LINEAR;
    br.colors.clear();
    br.stops.clear();
    br.start = v2(sx, sy);
    br.end = v2(ex, ey);
}

void canvas::set_radial_gradient(brush_type type, float sx, float sy, float sr,
                                 float ex, float ey, float er)
{
    if (sr < 0.0f || er < 0.0f) return;              // keeps the previous brush
    brush_state &br = pick(self, type);
    br.type = CB200_BRUSH_RADIAL;
    br.colors.clear();
    br.stops.clear();
    br.start = v2(sx, sy);
    br.end = v2(ex, ey);
    br.start_radius = sr;
    br.end_radius = er;
}

void canvas::add_color_stop(brush_type type, float offset, float r, float g, float b, float a)
{
    brush_state &peek = type == fill_style ? self->fill : self->stroke;
    if ((peek.type != CB200_BRUSH_LINEAR && peek.type != CB200_BRUSH_RADIAL) ||
        offset < 0.0f || 1.0f < offset)
        return;
    brush_state &br = pick(self, type);
    // later stops with an equal offset go after earlier ones (hpp:2831)
    size_t at = size_t(std::upper_bound(br.stops.begin(), br.stops.end(), offset) -
                       br.stops.begin());
    br.colors.insert(br.colors.begin() + ptrdiff_t(at), linear_of(r, g, b, a));
    br.stops.insert(br.stops.begin() + ptrdiff_t(at), offset);
}

static void load_pattern(brush_state &br, unsigned char const *image, int width,
                         int height, int stride, repetition_style repetition)
{
    br.type = CB200_BRUSH_PATTERN;
    br.colors.clear();
    br.texels.resize(size_t(width) * size_t(height) * 4);
    for (int y = 0; y < height; ++y)
        memcpy(&br.texels[size_t(y) * size_t(width) * 4], image + ptrdiff_t(y) * stride,
               size_t(width) * 4);
    // colors.front() is observable through clear_rectangle (hpp:3143-3149)
    br.colors.assign(1, texel_to_premultiplied_linear(&br.texels[0]));
    br.width = width;
    br.height = height;
    br.repetition = uint32_t(repetition);
}

void canvas::set_pattern(brush_type type, unsigned char const *image, int width,
                         int height, int stride, repetition_style repetition)
{
    if (!image || width <= 0 || height <= 0) return;
    load_pattern(pick(self, type), image, width, height, stride, repetition);
}

// ----------------------------------------------------------- path building ----

void canvas::begin_path()
{
    self->path.points.clear();
    self->path.subs.clear();
}

static void path_move(canvas::host_state *s, vec2 device_point)
{
    path_state &p = s->path;
    if (!p.subs.empty() && p.subs.back().count == 1) {   // bare move_to: overwrite
        p.points.back() = device_point;
        return;
    }
    path_state::sub fresh = { 1, false };
    p.points.push_back(device_point);
    p.subs.push_back(fresh);
}

static void path_line(canvas::host_state *s, vec2 device_point)
{
    path_state &p = s->path;
    vec2 from = p.points.back();
    vec2 d = device_point - from;
    if (dot(d, d) == 0.0f) return;                       // zero-length: dropped
    p.points.push_back(from);                            // line == cubic (a, a, b, b)
    p.points.push_back(device_point);
    p.points.push_back(device_point);
    p.subs.back().count += 3;
}

static void path_cubic(canvas::host_state *s, vec2 c1, vec2 c2, vec2 to)
{
    s->path.points.push_back(c1);
    s->path.points.push_back(c2);
    s->path.points.push_back(to);
    s->path.subs.back().count += 3;
}

void canvas::move_to(float x, float y) { path_move(self, apply(self->forward, v2(x, y))); }

void canvas::line_to(float x, float y)
{
    vec2 p = apply(self->forward, v2(x, y));
    if (self->path.subs.empty()) { path_move(self, p); return; }
    path_line(self, p);
}

void canvas::close_path()
{
    path_state &p = self->path;
    if (p.subs.empty()) return;
    vec2 first = p.points[p.points.size() - p.subs.back().count];   // already device space
    path_line(self, first);
    p.subs.back().closed = true;
    path_move(self, first);                              // next subpath starts here
}

void canvas::quadratic_curve_to(float cx, float cy, float x, float y)
{
    if (self->path.subs.empty()) move_to(cx, cy);
    vec2 from = self->path.points.back();
    vec2 ctrl = apply(self->forward, v2(cx, cy));
    vec2 to = apply(self->forward, v2(x, y));
    path_cubic(self, mix(from, ctrl, 2.0f / 3.0f), mix(to, ctrl, 2.0f / 3.0f), to);
}

void canvas::bezier_curve_to(float c1x, float c1y, float c2x, float c2y, float x, float y)
{
    if (self->path.subs.empty()) move_to(c1x, c1y);
    path_cubic(self, apply(self->forward, v2(c1x, c1y)), apply(self->forward, v2(c2x, c2y)),
               apply(self->forward, v2(x, y)));
}

void canvas::arc_to(float vx, float vy, float x, float y, float radius)
{
    if (radius < 0.0f || singular(self->forward)) return;
    if (self->path.subs.empty()) move_to(vx, vy);
    vec2 from = apply(self->inverse, self->path.points.back());    // back to user space
    vec2 corner = v2(vx, vy);
    vec2 u1 = unit(from - corner), u2 = unit(v2(x, y) - corner);
    float sine = fabsf(dot(perp(u1), u2));
    if (sine < 1.0e-4f) { line_to(vx, vy); return; }               // collinear
    vec2 to_center = (radius / sine) * (u1 + u2);
    vec2 center = corner + to_center;
    vec2 t1 = dot(to_center, u1) * u1 - to_center;                 // tangent points - center
    vec2 t2 = dot(to_center, u2) * u2 - to_center;
    float a1 = atan2f(t1.y, t1.x), a2 = atan2f(t2.y, t2.x);
    bool ccw = (int(floorf((a2 - a1) / 3.14159265f)) & 1) != 0;
    arc(center.x, center.y, radius, a1, a2, ccw);
}

void canvas::arc(float x, float y, float radius, float a0, float a1, bool counter_clockwise)
{
    if (radius < 0.0f) return;
    const float tau = 6.28318531f;
    float dir = counter_clockwise ? -1.0f : 1.0f;
    float from = fmodf(a0, tau);
    float sweep = fmodf(a1, tau) - from;
    if ((a1 - a0) * dir >= tau) sweep = tau * dir;       // full circle or more
    else if (sweep * dir < 0.0f) sweep += tau * dir;
    vec2 r0 = radius * v2(cosf(from), sinf(from));
    line_to(x + r0.x, y + r0.y);
    if (sweep == 0.0f) return;
    int pieces = int(std::max(1.0f, roundf(16.0f / tau * sweep * dir)));
    float step = sweep / float(pieces);
    float k = 4.0f / 3.0f * tanf(0.25f * step);          // cubic arc handle length
    for (int i = 0; i < pieces; ++i) {
        float ang = from + float(i + 1) * step;
        vec2 r1 = radius * v2(cosf(ang), sinf(ang));
        vec2 p0 = v2(x, y) + r0, p1 = v2(x, y) + r1;
        vec2 h0 = p0 + k * perp(r0), h1 = p1 - k * perp(r1);
        bezier_curve_to(h0.x, h0.y, h1.x, h1.y, p1.x, p1.y);
        r0 = r1;
    }
}

void canvas::rectangle(float x, float y, float w, float h)
{
    move_to(x, y);
    line_to(x + w, y);
    line_to(x + w, y + h);
    line_to(x, y + h);
    close_path();
}

// ------------------------------------------------------- frame builder ----

static cb200_glyph_atlas atlas_snapshot(glyph_cache &cache);   // defined with glyph_cache below

void canvas::host_state::reset_frame()
{
    draws.clear(); subpaths.clear(); points.clear(); brushes.clear();
    colors.clear(); stops.clear(); dashes.clear(); images.clear(); texels.clear();
    glyphs.clear(); frame_atlases.clear(); n_glyph_points = 0;
    cached_brush_serial[0] = cached_brush_serial[1] = cached_brush_serial[2] = 0;
}

void canvas::host_state::flush()
{
    if (draws.empty()) return;
    cb200_frame f;
    memset(&f, 0, sizeof f);
    f.draws = draws.data();       f.n_draws = uint32_t(draws.size());
    f.subpaths = subpaths.data(); f.n_subpaths = uint32_t(subpaths.size());
    f.points = points.data();     f.n_points = uint32_t(points.size() / 2);
    f.brushes = brushes.data();   f.n_brushes = uint32_t(brushes.size());
    f.colors = colors.data();     f.stops = stops.data();
    f.n_colors = uint32_t(stops.size());
    f.dashes = dashes.data();     f.n_dashes = uint32_t(dashes.size());
    f.images = images.data();     f.n_images = uint32_t(images.size());
    f.texels = texels.data();     f.texel_bytes = texels.size();
    std::vector<cb200_glyph_atlas> atlas_views;            // snapshots: everything the queued instances refer to
    for (size_t i = 0; i < frame_atlases.size(); ++i) atlas_views.push_back(atlas_snapshot(*frame_atlases[i]));
    f.atlases = atlas_views.data(); f.n_atlases = uint32_t(atlas_views.size());
    f.glyphs = glyphs.data();       f.n_glyphs = uint32_t(glyphs.size());
    f.n_glyph_points = n_glyph_points;
    if (tap.frame)
        tap.frame(tap.user, &f);
    else if (device) {
        int rc = cb200_submit(device, &f);
        if (rc != CB200_OK) {
            // The canvas API has no error channel (hpp: every method returns void):
            // a device failure is fatal rather than silently wrong pixels.
            fprintf(stderr, "canvas_b200: cb200_submit failed (%d): %s\n", rc,
                    cb200_last_error());
            abort();
        }
    }
    ++frames_flushed;
    reset_frame();
    // Clip-mask planes are immutable device slots; every so often tell the back end which ones
    // are still reachable (current mask + save stack) so the rest can be freed.
    if (device && next_mask - masks_at_last_keep >= 16) {
        std::vector<uint32_t> alive;
        if (mask) alive.push_back(mask);
        for (size_t i = 0; i < saves.size(); ++i)
            if (saves[i].mask) alive.push_back(saves[i].mask);
        cb200_masks_keep(device, alive.data(), uint32_t(alive.size()));
        masks_at_last_keep = next_mask;
    }
}

// Pool a brush; `which` 0/1/2 = fill/stroke/image lets unchanged brushes be
// shared by consecutive draws of one frame.
static uint32_t pool_brush(canvas::host_state *s, const brush_state &b, int which, bool clamp)
{
    if (which >= 0 && s->cached_brush_serial[which] == b.serial && b.serial != 0)
        return s->cached_brush_index[which];
    cb200_brush out;
    memset(&out, 0, sizeof out);
    out.type = b.type;
    out.flags = clamp ? CB200_BRUSH_CLAMP : 0;
    out.first_color = uint32_t(s->stops.size());
    out.start[0] = b.start.x; out.start[1] = b.start.y;
    out.end[0] = b.end.x;     out.end[1] = b.end.y;
    out.start_radius = b.start_radius;
    out.end_radius = b.end_radius;
    out.repetition = b.repetition;
    if (b.type == CB200_BRUSH_PATTERN) {
        cb200_image img;
        img.texel_offset = s->texels.size();
        img.width = b.width;
        img.height = b.height;
        s->texels.insert(s->texels.end(), b.texels.begin(), b.texels.end());
        while (s->texels.size() & 15) s->texels.push_back(0);      // keep rows 16 B aligned
        out.image = uint32_t(s->images.size());
        out.n_colors = b.texels.empty() ? 0 : 1;
        s->images.push_back(img);
    } else {
        out.n_colors = uint32_t(b.colors.size());
        for (size_t i = 0; i < b.colors.size(); ++i) {
            const color4 &c = b.colors[i];
            s->colors.push_back(c.r); s->colors.push_back(c.g);
            s->colors.push_back(c.b); s->colors.push_back(c.a);
            s->stops.push_back(i < b.stops.size() ? b.stops[i] : 0.0f);
        }
    }
    uint32_t index = uint32_t(s->brushes.size());
    s->brushes.push_back(out);
    if (which >= 0) {
        s->cached_brush_serial[which] = b.serial;
        s->cached_brush_index[which] = index;
    }
    return index;
}

// Subpath builder for geometry that bypasses `path` (rectangles, images, text):
// a start point plus cubics; straight segments are (from, from, to, to)... stored
// as the three trailing points (from, to, to) exactly like line_to does.
struct outline_builder {
    canvas::host_state *s;
    uint32_t first_subpath;
    bool open = false;
    vec2 start = {0, 0}, last = {0, 0};
    uint32_t cubics = 0;
    uint32_t first_point = 0;

    explicit outline_builder(canvas::host_state *state)
        : s(state), first_subpath(uint32_t(state->subpaths.size())) {}
    void put(vec2 p) { s->points.push_back(p.x); s->points.push_back(p.y); }
    void begin(vec2 p)
    {
        first_point = uint32_t(s->points.size() / 2);
        put(p);
        start = last = p;
        cubics = 0;
        open = true;
    }
    void line(vec2 p) { put(last); put(p); put(p); last = p; ++cubics; }
    void cubic(vec2 c1, vec2 c2, vec2 p) { put(c1); put(c2); put(p); last = p; ++cubics; }
    void end(bool closed)
    {
        if (!open) return;
        cb200_subpath sp = { first_point, cubics, closed ? 1u : 0u, 0u };
        s->subpaths.push_back(sp);
        open = false;
    }
    uint32_t count() const { return uint32_t(s->subpaths.size()) - first_subpath; }
};

static void copy_affine(float out[6], const affine &m)
{
    out[0] = m.a; out[1] = m.b; out[2] = m.c; out[3] = m.d; out[4] = m.e; out[5] = m.f;
}

static void queue_draw(canvas &cv, canvas::host_state *s, uint32_t kind, uint32_t brush,
                       uint32_t first_subpath, uint32_t n_subpaths)
{
    cb200_draw d;
    memset(&d, 0, sizeof d);
    d.kind = kind;
    d.op = uint32_t(cv.global_composite_operation);
    d.first_subpath = first_subpath;
    d.n_subpaths = n_subpaths;
    d.brush = brush;
    d.mask_src = s->mask;
    d.cap = uint32_t(cv.line_cap);
    d.join = uint32_t(cv.line_join);
    d.global_alpha = s->global_alpha;
    d.line_width = s->line_width;
    d.miter_limit = s->miter_limit;
    copy_affine(d.forward, s->forward);
    copy_affine(d.inverse, s->inverse);
    if (kind == CB200_STROKE && !s->dash.empty()) {
        d.first_dash = uint32_t(s->dashes.size());
        d.n_dash = uint32_t(s->dash.size());
        d.dash_offset = cv.line_dash_offset;
        s->dashes.insert(s->dashes.end(), s->dash.begin(), s->dash.end());
    }
    d.shadow_color[0] = s->shadow_color.r; d.shadow_color[1] = s->shadow_color.g;
    d.shadow_color[2] = s->shadow_color.b; d.shadow_color[3] = s->shadow_color.a;
    d.shadow_offset_x = cv.shadow_offset_x;
    d.shadow_offset_y = cv.shadow_offset_y;
    d.shadow_blur = s->shadow_blur;
    if (kind == CB200_CLIP) {
        d.mask_dst = s->next_mask++;
        s->mask = d.mask_dst;
    }
    s->draws.push_back(d);
    if (s->draws.size() >= s->max_queued_draws) s->flush();
}

// Copy the current path into the frame pools (bare move_to subpaths carry no
// geometry for fill/stroke/clip and are dropped).
static uint32_t pool_path(canvas::host_state *s, uint32_t &n_subpaths)
{
    uint32_t first = uint32_t(s->subpaths.size());
    size_t at = 0;
    for (size_t i = 0; i < s->path.subs.size(); ++i) {
        const path_state::sub &sub = s->path.subs[i];
        if (sub.count >= 4) {
            cb200_subpath sp = { uint32_t(s->points.size() / 2), (sub.count - 1) / 3,
                                 sub.closed ? 1u : 0u, 0u };
            for (size_t k = 0; k < sub.count; ++k) {
                s->points.push_back(s->path.points[at + k].x);
                s->points.push_back(s->path.points[at + k].y);
            }
            s->subpaths.push_back(sp);
        }
        at += sub.count;
    }
    n_subpaths = uint32_t(s->subpaths.size()) - first;
    return first;
}

// ------------------------------------------------------------ draw calls ----

void canvas::fill()
{
    if (singular(self->forward)) return;                 // render_main's early out
    uint32_t n, first = pool_path(self, n);
    queue_draw(*this, self, CB200_FILL, pool_brush(self, self->fill, 0, false), first, n);
}

void canvas::stroke()
{
    if (singular(self->forward)) return;
    uint32_t n, first = pool_path(self, n);
    queue_draw(*this, self, CB200_STROKE, pool_brush(self, self->stroke, 1, false), first, n);
}

void canvas::clip()
{
    uint32_t n, first = pool_path(self, n);
    queue_draw(*this, self, CB200_CLIP, 0, first, n);
}

static void rectangle_outline(outline_builder &ob, const affine &m, float x, float y,
                              float w, float h, bool repeat_first)
{
    ob.begin(apply(m, v2(x, y)));
    ob.line(apply(m, v2(x + w, y)));
    ob.line(apply(m, v2(x + w, y + h)));
    ob.line(apply(m, v2(x, y + h)));
    if (repeat_first) ob.line(apply(m, v2(x, y)));
    ob.end(true);
}

void canvas::fill_rectangle(float x, float y, float w, float h)
{
    if (w == 0.0f || h == 0.0f || singular(self->forward)) return;
    outline_builder ob(self);
    rectangle_outline(ob, self->forward, x, y, w, h, false);
    queue_draw(*this, self, CB200_FILL, pool_brush(self, self->fill, 0, false),
               ob.first_subpath, ob.count());
}

void canvas::stroke_rectangle(float x, float y, float w, float h)
{
    if ((w == 0.0f && h == 0.0f) || singular(self->forward)) return;
    outline_builder ob(self);
    if (w == 0.0f || h == 0.0f) {                        // degenerate: one open segment
        ob.begin(apply(self->forward, v2(x, y)));
        ob.line(apply(self->forward, v2(x + w, y + h)));
        ob.end(false);
    } else
        rectangle_outline(ob, self->forward, x, y, w, h, true);
    queue_draw(*this, self, CB200_STROKE, pool_brush(self, self->stroke, 1, false),
               ob.first_subpath, ob.count());
}

void canvas::clear_rectangle(float x, float y, float w, float h)
{
    if (w == 0.0f || h == 0.0f || singular(self->forward)) return;
    // destination_out with the fill brush forced to "solid" WITHOUT touching its
    // colour list (hpp:3143-3149): whatever colors.front() holds does the erasing.
    brush_state eraser;
    eraser.type = CB200_BRUSH_COLOR;
    if (!self->fill.colors.empty()) eraser.colors.assign(1, self->fill.colors.front());
    composite_operation keep_op = global_composite_operation;
    float keep_alpha = self->global_alpha, keep_shadow = self->shadow_color.a;
    global_composite_operation = destination_out;
    self->global_alpha = 1.0f;
    self->shadow_color.a = 0.0f;
    outline_builder ob(self);
    rectangle_outline(ob, self->forward, x, y, w, h, false);
    queue_draw(*this, self, CB200_FILL, pool_brush(self, eraser, -1, false),
               ob.first_subpath, ob.count());
    self->shadow_color.a = keep_shadow;
    self->global_alpha = keep_alpha;
    global_composite_operation = keep_op;
}

void canvas::draw_image(unsigned char const *image, int width, int height, int stride,
                        float x, float y, float to_w, float to_h)
{
    if (!image || width <= 0 || height <= 0 || to_w == 0.0f || to_h == 0.0f) return;
    self->image.serial = ++self->serial_counter;
    load_pattern(self->image, image, width, height, stride, repeat);
    outline_builder ob(self);
    rectangle_outline(ob, self->forward, x, y, to_w, to_h, false);
    affine keep_f = self->forward, keep_i = self->inverse;
    translate(x + std::min(0.0f, to_w), y + std::min(0.0f, to_h));   // brush space = texels
    scale(fabsf(to_w) / float(width), fabsf(to_h) / float(height));
    if (!singular(self->forward))
        queue_draw(*this, self, CB200_FILL, pool_brush(self, self->image, 2, true),
                   ob.first_subpath, ob.count());
    self->forward = keep_f;
    self->inverse = keep_i;
}

// The current path flattened for hit testing (path_to_lines( false ), hpp:3105 / 1495-1524) as
// edges (from, to), every subpath's closing edge included -- with the same routine the device's
// K1 uses (geom.cuh), in the reference's point order.
void flattened_path_edges(const canvas::host_state *self, std::vector<float> &edges)
{
    edges.clear();
    size_t at = 0;
    std::vector<vec2> poly;
    for (size_t i = 0; i < self->path.subs.size(); ++i) {
        const path_state::sub &sub = self->path.subs[i];
        poly.clear();
        poly.push_back(self->path.points[at]);
        for (size_t k = 1; k + 2 < sub.count + 0u; k += 3) {
            struct vec_sink { std::vector<vec2> *v; void put(vec2 p) { v->push_back(p); } } sink = { &poly };
            vec2 from = self->path.points[at + k - 1];
            cb200::flatten_cubic(from, self->path.points[at + k], self->path.points[at + k + 1],
                                 self->path.points[at + k + 2], -1.0f, sink);
        }
        at += sub.count;
        for (size_t k = 0; k < poly.size(); ++k) {
            vec2 a = poly[k], b = poly[k + 1 < poly.size() ? k + 1 : 0];
            edges.push_back(a.x); edges.push_back(a.y); edges.push_back(b.x); edges.push_back(b.y);
        }
    }
}

bool canvas::is_point_in_path(float x, float y)
{
    // Synchronous single-point query: count signed crossings on the host (hpp:3101-3132).  Many
    // points at once go to the device instead (cv_points_in_path / cb200_hit_test).
    std::vector<float> edges;
    flattened_path_edges(self, edges);
    int winding = 0;
    for (size_t k = 0; k < edges.size(); k += 4) {
        vec2 a = v2(edges[k], edges[k + 1]), b = v2(edges[k + 2], edges[k + 3]);
        if ((a.y < y && y <= b.y) || (b.y < y && y <= a.y)) {
            float side = dot(perp(b - a), v2(x, y) - a);
            if (side == 0.0f) return true;               // on an edge
            winding += side > 0.0f ? 1 : -1;
        } else if (a.y == y && y == b.y &&
                   ((a.x <= x && x <= b.x) || (b.x <= x && x <= a.x)))
            return true;                                 // on a horizontal edge
    }
    return winding != 0;
}

// ------------------------------------------------------------------ text ----

namespace {
struct ttf {
    const std::vector<uint8_t> &d;
    int u8(int i) const { return d[size_t(i)]; }
    int s8(int i) const { return int(int8_t(d[size_t(i)])); }
    int u16(int i) const { return d[size_t(i)] << 8 | d[size_t(i) + 1]; }
    int s16(int i) const { return int(int16_t(u16(i))); }
    int s32(int i) const
    {
        return int(uint32_t(d[size_t(i)]) << 24 | uint32_t(d[size_t(i) + 1]) << 16 |
                   uint32_t(d[size_t(i) + 2]) << 8 | uint32_t(d[size_t(i) + 3]));
    }
};
}

// ---- glyph outline cache (include/canvas_b200.h, "glyph outline cache") ----------------
//
// One cache per distinct font (shared by content between canvases).  A glyph is parsed on first
// use into a transform-independent outline -- the contour walk of add_glyph (hpp:1533-1696) run
// once with SYMBOLIC points -- and appended to the atlas the back end mirrors in device memory.
// Arrays grow by generations that are never freed, so a snapshot handed to a frame (and any deep
// copy of that frame) stays valid while other canvases keep adding glyphs.
template <class T>
struct stable_array {
    T *data = nullptr;
    size_t size = 0, cap = 0;
    std::vector<std::unique_ptr<T[]> > generations;
    void push(const T &v)
    {
        if (size == cap) {
            size_t grown = cap ? cap * 2 : 256;
            std::unique_ptr<T[]> g(new T[grown]);
            for (size_t i = 0; i < size; ++i) g[i] = data[i];
            data = g.get();
            cap = grown;
            generations.push_back(std::move(g));
        }
        data[size++] = v;
    }
};

struct glyph_cache {
    struct component { int glyph; float a, b, c, d, e, f; };
    struct contour { uint32_t out_first, n_cubics; };
    struct entry {
        enum kind_t { EMPTY, SIMPLE, COMPOSITE, HOST_ONLY } kind = EMPTY;
        uint32_t outline = 0, out_points = 0;
        std::vector<contour> contours;       // SIMPLE: one closed subpath each
        std::vector<component> parts;        // COMPOSITE: child glyph + its 2x3 placement
    };

    std::mutex lock;
    uint64_t id = 0;
    std::vector<uint8_t> key;                // the font bytes (font_state::data) this cache is for
    stable_array<cb200_glyph_outline> outlines;
    stable_array<cb200_glyph_seg> segs;
    stable_array<float> points;
    std::unordered_map<int, std::unique_ptr<entry> > glyphs;   // entries are immutable once built

    const entry &lookup(const font_state &f, int glyph)
    {
        std::lock_guard<std::mutex> hold(lock);
        std::unique_ptr<entry> &slot = glyphs[glyph];
        if (!slot) { slot.reset(new entry); build(f, glyph, *slot); }
        return *slot;
    }
    cb200_glyph_atlas snapshot()
    {
        std::lock_guard<std::mutex> hold(lock);
        cb200_glyph_atlas a;
        memset(&a, 0, sizeof a);
        a.id = id;
        a.outlines = outlines.data; a.n_outlines = uint32_t(outlines.size);
        a.segs = segs.data;         a.n_segs = uint32_t(segs.size);
        a.points = points.data;     a.n_points = uint32_t(points.size / 2);
        return a;
    }
    static std::shared_ptr<glyph_cache> for_font(const std::vector<uint8_t> &data)
    {
        static std::mutex registry_lock;
        static std::multimap<uint64_t, std::shared_ptr<glyph_cache> > registry;   // kept for the life of the process
        static std::atomic<uint64_t> next_id(1);
        uint64_t h = 1469598103934665603ull;                                       // FNV-1a
        for (size_t i = 0; i < data.size(); ++i) h = (h ^ data[i]) * 1099511628211ull;
        std::lock_guard<std::mutex> hold(registry_lock);
        auto range = registry.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (it->second->key == data) return it->second;
        std::shared_ptr<glyph_cache> fresh(new glyph_cache);
        fresh->id = next_id++;
        fresh->key = data;
        registry.insert(std::make_pair(h, fresh));
        return fresh;
    }

private:
    struct ep { uint16_t a, b; };            // P[a], or mix(P[a], P[b], 0.5) when a != b
    void build(const font_state &f, int glyph, entry &e);
};

static cb200_glyph_atlas atlas_snapshot(glyph_cache &cache) { return cache.snapshot(); }

// The contour walk of lower_glyph() below with point INDICES in place of transformed points.
void glyph_cache::build(const font_state &f, int glyph, entry &e)
{
    ttf r = { f.data };
    bool long_loca = r.u16(f.head + 50) != 0;
    int at = f.glyf + (long_loca ? r.s32(f.loca + glyph * 4) : r.u16(f.loca + glyph * 2) * 2);
    int stop = f.glyf + (long_loca ? r.s32(f.loca + glyph * 4 + 4)
                                   : r.u16(f.loca + glyph * 2 + 2) * 2);
    if (at == stop) return;                              // empty glyph (space)
    int contours = r.s16(at);
    if (contours < 0) {                                  // composite glyph: the parts and their placement
        e.kind = entry::COMPOSITE;
        at += 10;
        for (;;) {
            int flags = r.u16(at), part = r.u16(at + 2);
            if (!(flags & 2)) return;                    // point matching unsupported
            component c;
            c.glyph = part;
            c.e = float(flags & 1 ? r.s16(at + 4) : r.s8(at + 4));
            c.f = float(flags & 1 ? r.s16(at + 6) : r.s8(at + 5));
            at += flags & 1 ? 8 : 6;
            c.a = flags & 200 ? float(r.s16(at)) / 16384.0f : 1.0f;
            c.b = flags & 128 ? float(r.s16(at + 2)) / 16384.0f : 0.0f;
            c.c = flags & 128 ? float(r.s16(at + 4)) / 16384.0f : 0.0f;
            c.d = flags & 8 ? c.a : flags & 64 ? float(r.s16(at + 2)) / 16384.0f :
                  flags & 128 ? float(r.s16(at + 6)) / 16384.0f : 1.0f;
            at += flags & 8 ? 2 : flags & 64 ? 4 : flags & 128 ? 8 : 0;
            e.parts.push_back(c);
            if (!(flags & 32)) return;                   // no more components
        }
    }
    int hmetrics = r.u16(f.hhea + 34);
    int lsb = glyph < hmetrics ? r.s16(f.hmtx + glyph * 4 + 2)
                               : r.s16(f.hmtx + hmetrics * 2 + glyph * 2);
    int x_min = r.s16(at + 2);
    int n_points = contours ? r.u16(at + 8 + contours * 2) + 1 : 0;
    e.kind = entry::HOST_ONLY;                           // until the whole outline is known to be expressible
    if (contours == 0 || n_points > 65535) return;
    int flag_at = at + 12 + contours * 2 + r.u16(at + 10 + contours * 2);
    int flag_bytes = 0, x_bytes = 0;
    for (int i = 0; i < n_points;) {
        int fl = r.u8(flag_at + flag_bytes++);
        int rep = fl & 8 ? r.u8(flag_at + flag_bytes++) + 1 : 1;
        x_bytes += rep * (fl & 2 ? 1 : fl & 16 ? 0 : 2);
        i += rep;
    }
    int x_at = flag_at + flag_bytes, y_at = x_at + x_bytes;
    int x = lsb - x_min, y = 0, fl = 0, rep = 0, i = 0;
    std::vector<float> pts;
    std::vector<cb200_glyph_seg> pieces;
    uint32_t out = 0;                                    // next output point of the instance
    ep start = {0, 0}, last = {0, 0};
    bool first_piece = false;
    auto begin = [&](ep p) { start = last = p; first_piece = true; ++out; };
    auto piece = [&](uint16_t ctrl, ep to, bool straight) {
        cb200_glyph_seg sg = { last.a, last.b, to.a, to.b, ctrl,
                               uint16_t((straight ? CB200_SEG_LINE : 0) | (first_piece ? CB200_SEG_FIRST : 0)), out };
        pieces.push_back(sg);
        out += 3;
        last = to;
        first_piece = false;
    };
    for (int c = 0; c < contours; ++c) {
        int first_index = i, last_index = r.u16(at + 10 + c * 2);
        if (last_index < first_index || last_index >= n_points) return;     // malformed: host lowering keeps its behaviour
        uint16_t begin_idx = 0, prev_idx = 0;
        bool begin_on = false, prev_on = false, started = false;
        contour ct = { out, 0 };
        size_t pieces_before = pieces.size();
        for (; i <= last_index; ++i) {
            if (rep) --rep;
            else {
                fl = r.u8(flag_at++);
                if (fl & 8) rep = r.u8(flag_at++);
            }
            if (fl & 2) x += r.u8(x_at) * (fl & 16 ? 1 : -1);
            else if (!(fl & 16)) x += r.s16(x_at);
            if (fl & 4) y += r.u8(y_at) * (fl & 32 ? 1 : -1);
            else if (!(fl & 32)) y += r.s16(y_at);
            x_at += fl & 2 ? 1 : fl & 16 ? 0 : 2;
            y_at += fl & 4 ? 1 : fl & 32 ? 0 : 2;
            pts.push_back(float(x));
            pts.push_back(float(y));
            uint16_t idx = uint16_t(i);
            bool on = (fl & 1) != 0;
            if (i == first_index) {
                begin_idx = idx;
                begin_on = on;
                if (on) { ep p = { idx, idx }; begin(p); started = true; }
            } else {
                ep to = { on ? idx : prev_idx, idx };                // implied on-curve midpoint
                if (!started) { begin(to); started = true; }
                else if (prev_on && on) piece(0, to, true);
                else if (!prev_on || on) piece(prev_idx, to, false);  // quadratic around the previous point
            }
            prev_idx = idx;
            prev_on = on;
        }
        if (!started) { ep p = { begin_idx, begin_idx }; begin(p); }
        if (begin_on != prev_on) piece(prev_on ? begin_idx : prev_idx, start, false);
        else if (!begin_on && !prev_on) {
            ep half = { begin_idx, prev_idx };
            piece(prev_idx, half, false);
            piece(begin_idx, start, false);
        }
        piece(0, start, true);                                       // explicit closing point
        ct.n_cubics = uint32_t(pieces.size() - pieces_before);
        e.contours.push_back(ct);
    }
    cb200_glyph_outline o = { uint32_t(points.size / 2), uint32_t(n_points), uint32_t(segs.size),
                              uint32_t(pieces.size()), uint32_t(contours), out };
    for (size_t k = 0; k < pts.size(); ++k) points.push(pts[k]);
    for (size_t k = 0; k < pieces.size(); ++k) segs.push(pieces[k]);
    e.outline = uint32_t(outlines.size);
    e.out_points = out;
    outlines.push(o);
    e.kind = entry::SIMPLE;
}

bool canvas::set_font(unsigned char const *font, int bytes, float size)
{
    font_state &f = self->face;
    if (font && bytes) {
        f = font_state();
        if (bytes < 6) return false;
        uint32_t version = uint32_t(font[0]) << 24 | uint32_t(font[1]) << 16 |
                           uint32_t(font[2]) << 8 | uint32_t(font[3]);
        int tables = font[4] << 8 | font[5];
        if ((version != 0x00010000u && version != 0x74727565u) || bytes < tables * 16 + 12)
            return false;
        f.data.assign(font, font + tables * 16 + 12);
        static const struct { uint32_t tag; int font_state::*slot; } wanted[] = {
            { 0x636d6170u, &font_state::cmap }, { 0x676c7966u, &font_state::glyf },
            { 0x68656164u, &font_state::head }, { 0x68686561u, &font_state::hhea },
            { 0x686d7478u, &font_state::hmtx }, { 0x6c6f6361u, &font_state::loca },
            { 0x6d617870u, &font_state::maxp }, { 0x4f532f32u, &font_state::os_2 } };
        for (int t = 0; t < tables; ++t) {
            ttf dir = { f.data };
            uint32_t tag = uint32_t(dir.s32(t * 16 + 12));
            int offset = dir.s32(t * 16 + 20), span = dir.s32(t * 16 + 24);
            if (bytes < offset + span) { f.data.clear(); return false; }
            for (size_t w = 0; w < sizeof wanted / sizeof wanted[0]; ++w)
                if (wanted[w].tag == tag) {
                    f.*(wanted[w].slot) = int(f.data.size());
                    f.data.insert(f.data.end(), font + offset, font + offset + span);
                    break;
                }
        }
        if (!f.cmap || !f.glyf || !f.head || !f.hhea || !f.hmtx || !f.loca || !f.maxp ||
            !f.os_2) {
            f.data.clear();
            return false;
        }
        f.cache = glyph_cache::for_font(f.data);
    }
    if (f.data.empty()) return false;
    ttf r = { f.data };
    f.scale = size / float(r.u16(f.head + 18));          // size / unitsPerEm
    return true;
}

// UTF-8 -> code point -> glyph id through cmap format 12, 4 or 0 (hpp:1709-1784).
static int next_glyph(const font_state &f, char const *text, int &at)
{
    ttf r = { f.data };
    int lead = text[at];
    int len = (lead & 0x80) == 0x00 ? 1 : (lead & 0xe0) == 0xc0 ? 2 :
              (lead & 0xf0) == 0xe0 ? 3 : (lead & 0xf8) == 0xf0 ? 4 : 0;
    static const int lead_mask[] = { 0x0, 0x7f, 0x1f, 0x0f, 0x07 };
    int cp = len ? lead & lead_mask[len] : 0xfffd;
    ++at;
    while (--len > 0) {
        if ((text[at] & 0xc0) != 0x80) { cp = 0xfffd; break; }    // resync on the bad byte
        cp = cp << 6 | (text[at++] & 0x3f);
    }
    if (cp == '\t' || cp == '\v' || cp == '\f' || cp == '\r' || cp == '\n') cp = ' ';
    int fmt12 = 0, fmt4 = 0, fmt0 = 0;
    int n = r.u16(f.cmap + 2);
    for (int t = 0; t < n; ++t) {
        int platform = r.u16(f.cmap + t * 8 + 4), encoding = r.u16(f.cmap + t * 8 + 6);
        int sub = f.cmap + r.s32(f.cmap + t * 8 + 8);
        int format = r.u16(sub);
        if (platform == 3 && encoding == 10 && format == 12) fmt12 = sub;
        else if (platform == 3 && encoding == 1 && format == 4) fmt4 = sub;
        else if (format == 0) fmt0 = sub;
    }
    if (fmt12) {
        int groups = r.s32(fmt12 + 12);
        for (int g = 0; g < groups; ++g) {
            int lo = r.s32(fmt12 + 16 + g * 12), hi = r.s32(fmt12 + 20 + g * 12);
            if (lo <= cp && cp <= hi) return cp - lo + r.s32(fmt12 + 24 + g * 12);
        }
    } else if (fmt4) {
        int seg2 = r.u16(fmt4 + 6);                      // segCountX2
        int ends = fmt4 + 14, starts = ends + 2 + seg2, deltas = starts + seg2,
            ranges = deltas + seg2;
        for (int k = 0; k < seg2; k += 2) {
            int lo = r.u16(starts + k), hi = r.u16(ends + k);
            if (lo <= cp && cp <= hi) {
                int range = r.u16(ranges + k);
                return range ? r.u16(ranges + k + (cp - lo) * 2 + range)
                             : (cp + r.s16(deltas + k)) & 0xffff;
            }
        }
    } else if (fmt0 && 0 <= cp && cp < 256)
        return r.u8(fmt0 + 6 + cp);
    return 0;
}

static int advance_of(const font_state &f, int glyph)
{
    ttf r = { f.data };
    int hmetrics = r.u16(f.hhea + 34);
    return r.u16(f.hmtx + std::min(glyph, hmetrics - 1) * 4);
}

float canvas::measure_text(char const *text)
{
    if (self->face.data.empty() || !text) return 0.0f;
    int width = 0;
    for (int at = 0; text[at];) width += advance_of(self->face, next_glyph(self->face, text, at));
    return float(width) * self->face.scale;
}

// One glyph -> subpaths of cubics under the matrix `m` (font units -> device).
// Same contour walk as hpp:1533-1696, but the quadratic pieces are kept as
// (degree-elevated) cubics for the device to flatten instead of being
// tessellated here.
static void lower_glyph(outline_builder &ob, const font_state &f, int glyph, const affine &m)
{
    ttf r = { f.data };
    bool long_loca = r.u16(f.head + 50) != 0;
    int at = f.glyf + (long_loca ? r.s32(f.loca + glyph * 4) : r.u16(f.loca + glyph * 2) * 2);
    int stop = f.glyf + (long_loca ? r.s32(f.loca + glyph * 4 + 4)
                                   : r.u16(f.loca + glyph * 2 + 2) * 2);
    if (at == stop) return;                              // empty glyph (space)
    int contours = r.s16(at);
    if (contours < 0) {                                  // composite glyph
        at += 10;
        for (;;) {
            int flags = r.u16(at), part = r.u16(at + 2);
            if (!(flags & 2)) return;                    // point matching unsupported
            float e = float(flags & 1 ? r.s16(at + 4) : r.s8(at + 4));
            float ff = float(flags & 1 ? r.s16(at + 6) : r.s8(at + 5));
            at += flags & 1 ? 8 : 6;
            float a = flags & 200 ? float(r.s16(at)) / 16384.0f : 1.0f;
            float b = flags & 128 ? float(r.s16(at + 2)) / 16384.0f : 0.0f;
            float c = flags & 128 ? float(r.s16(at + 4)) / 16384.0f : 0.0f;
            float d = flags & 8 ? a : flags & 64 ? float(r.s16(at + 2)) / 16384.0f :
                      flags & 128 ? float(r.s16(at + 6)) / 16384.0f : 1.0f;
            at += flags & 8 ? 2 : flags & 64 ? 4 : flags & 128 ? 8 : 0;
            affine child = { m.a * a + m.c * b, m.b * a + m.d * b,
                             m.a * c + m.c * d, m.b * c + m.d * d,
                             m.a * e + m.c * ff + m.e, m.b * e + m.d * ff + m.f };
            lower_glyph(ob, f, part, child);
            if (!(flags & 32)) return;                   // no more components
        }
    }
    int hmetrics = r.u16(f.hhea + 34);
    int lsb = glyph < hmetrics ? r.s16(f.hmtx + glyph * 4 + 2)
                               : r.s16(f.hmtx + hmetrics * 2 + glyph * 2);
    int x_min = r.s16(at + 2);
    int n_points = r.u16(at + 8 + contours * 2) + 1;
    int flag_at = at + 12 + contours * 2 + r.u16(at + 10 + contours * 2);
    int flag_bytes = 0, x_bytes = 0;
    for (int i = 0; i < n_points;) {
        int fl = r.u8(flag_at + flag_bytes++);
        int rep = fl & 8 ? r.u8(flag_at + flag_bytes++) + 1 : 1;
        x_bytes += rep * (fl & 2 ? 1 : fl & 16 ? 0 : 2);
        i += rep;
    }
    int x_at = flag_at + flag_bytes, y_at = x_at + x_bytes;
    int x = lsb - x_min, y = 0, fl = 0, rep = 0, i = 0;
    for (int c = 0; c < contours; ++c) {
        int first_index = i, last_index = r.u16(at + 10 + c * 2);
        vec2 begin_pt = v2(0, 0), prev_pt = v2(0, 0);
        bool begin_on = false, prev_on = false, started = false;
        for (; i <= last_index; ++i) {
            if (rep) --rep;
            else {
                fl = r.u8(flag_at++);
                if (fl & 8) rep = r.u8(flag_at++);
            }
            if (fl & 2) x += r.u8(x_at) * (fl & 16 ? 1 : -1);
            else if (!(fl & 16)) x += r.s16(x_at);
            if (fl & 4) y += r.u8(y_at) * (fl & 32 ? 1 : -1);
            else if (!(fl & 32)) y += r.s16(y_at);
            x_at += fl & 2 ? 1 : fl & 16 ? 0 : 2;
            y_at += fl & 4 ? 1 : fl & 32 ? 0 : 2;
            vec2 pt = apply(m, v2(float(x), float(y)));
            bool on = (fl & 1) != 0;
            if (i == first_index) {
                begin_pt = pt;
                begin_on = on;
                if (on) { ob.begin(pt); started = true; }
            } else {
                vec2 to = on ? pt : mix(prev_pt, pt, 0.5f);      // implied on-curve midpoint
                if (!started) { ob.begin(to); started = true; }
                else if (prev_on && on) ob.line(to);
                else if (!prev_on || on)                         // quadratic around prev_pt
                    ob.cubic(mix(ob.last, prev_pt, 2.0f / 3.0f), mix(to, prev_pt, 2.0f / 3.0f), to);
            }
            prev_pt = pt;
            prev_on = on;
        }
        if (!started) { ob.begin(begin_pt); }                    // defensive: 1-point contour
        if (begin_on != prev_on) {
            vec2 ctrl = prev_on ? begin_pt : prev_pt;
            vec2 to = ob.start;
            ob.cubic(mix(ob.last, ctrl, 2.0f / 3.0f), mix(to, ctrl, 2.0f / 3.0f), to);
        } else if (!begin_on && !prev_on) {
            vec2 from = ob.last, to = ob.start;
            vec2 half = mix(begin_pt, prev_pt, 0.5f);
            ob.cubic(mix(from, prev_pt, 2.0f / 3.0f), mix(half, prev_pt, 2.0f / 3.0f), half);
            ob.cubic(mix(half, begin_pt, 2.0f / 3.0f), mix(to, begin_pt, 2.0f / 3.0f), to);
        }
        ob.line(ob.start);                                       // explicit closing point
        ob.end(true);
    }
}

// One glyph as instance records: the cached outline(s) + this draw's matrix; the device writes the
// control points lower_glyph() would have produced (bit for bit) into the frame's point pool.
static void instance_glyph(canvas::host_state *s, outline_builder &ob, const font_state &f, int glyph,
                           const affine &m)
{
    const glyph_cache::entry &e = f.cache->lookup(f, glyph);
    if (e.kind == glyph_cache::entry::EMPTY) return;
    if (e.kind == glyph_cache::entry::HOST_ONLY) { lower_glyph(ob, f, glyph, m); return; }
    if (e.kind == glyph_cache::entry::COMPOSITE) {
        for (size_t k = 0; k < e.parts.size(); ++k) {
            const glyph_cache::component &p = e.parts[k];
            affine child = { m.a * p.a + m.c * p.b, m.b * p.a + m.d * p.b,
                             m.a * p.c + m.c * p.d, m.b * p.c + m.d * p.d,
                             m.a * p.e + m.c * p.f + m.e, m.b * p.e + m.d * p.f + m.f };
            instance_glyph(s, ob, f, p.glyph, child);
        }
        return;
    }
    uint32_t atlas = 0;
    while (atlas < s->frame_atlases.size() && s->frame_atlases[atlas] != f.cache) ++atlas;
    if (atlas == s->frame_atlases.size()) s->frame_atlases.push_back(f.cache);
    cb200_glyph_inst gi = { atlas, e.outline, s->n_glyph_points, { m.a, m.b, m.c, m.d, m.e, m.f } };
    s->glyphs.push_back(gi);
    for (size_t c = 0; c < e.contours.size(); ++c) {
        cb200_subpath sp = { gi.first_point + e.contours[c].out_first, e.contours[c].n_cubics, 1u, 1u };
        s->subpaths.push_back(sp);
    }
    s->n_glyph_points += e.out_points;
}

// Text layout (hpp:1793-1846): alignment, baseline, max-width squeeze, then one
// font-units -> device matrix per glyph.
static void lower_text(canvas &cv, canvas::host_state *s, outline_builder &ob,
                       char const *text, float px, float py, float max_width)
{
    const font_state &f = s->face;
    if (f.data.empty() || !text || max_width <= 0.0f) return;
    ttf r = { f.data };
    float width = max_width == 1.0e30f && cv.text_align == leftward ? 0.0f : cv.measure_text(text);
    float squeeze = max_width / std::max(max_width, width);
    if (cv.text_align == rightward) px -= width * squeeze;
    else if (cv.text_align == center) px -= 0.5f * width * squeeze;
    float sx = f.scale * squeeze, sy = f.scale * 1.0f;
    float em = float(r.u16(f.head + 18));
    float ascender = float(r.s16(f.os_2 + 68)), descender = float(r.s16(f.os_2 + 70));
    float norm = f.scale * em / (ascender - descender);
    if (cv.text_baseline == top) py += ascender * norm;
    else if (cv.text_baseline == middle) py += (ascender + descender) * 0.5f * norm;
    else if (cv.text_baseline == bottom) py += descender * norm;
    else if (cv.text_baseline == hanging) py += 0.6f * f.scale * em;
    const affine &m = s->forward;
    int pen = 0;
    for (int at = 0; text[at];) {
        int glyph = next_glyph(f, text, at);
        float e = px + float(pen) * sx;
        affine g = { m.a * sx + m.c * 0.0f, m.b * sx + m.d * 0.0f,
                     m.a * 0.0f + m.c * -sy, m.b * 0.0f + m.d * -sy,
                     m.a * e + m.c * py + m.e, m.b * e + m.d * py + m.f };
        if (s->instanced_text && f.cache) instance_glyph(s, ob, f, glyph, g);
        else lower_glyph(ob, f, glyph, g);
        pen += advance_of(f, glyph);
    }
}

void canvas::fill_text(char const *text, float x, float y, float max_width)
{
    outline_builder ob(self);
    lower_text(*this, self, ob, text, x, y, max_width);
    if (singular(self->forward)) return;
    queue_draw(*this, self, CB200_FILL, pool_brush(self, self->fill, 0, false),
               ob.first_subpath, ob.count());
}

void canvas::stroke_text(char const *text, float x, float y, float max_width)
{
    outline_builder ob(self);
    lower_text(*this, self, ob, text, x, y, max_width);
    if (singular(self->forward)) return;
    queue_draw(*this, self, CB200_STROKE, pool_brush(self, self->stroke, 1, false),
               ob.first_subpath, ob.count());
}

// --------------------------------------------------------- pixels in/out ----

void canvas::get_image_data(unsigned char *image, int width, int height, int stride,
                            int x, int y)
{
    if (!image) return;
    self->flush();
    if (self->tap.read_rgba8)
        self->tap.read_rgba8(self->tap.user, image, width, height, stride, x, y);
    else if (self->device) {
        int rc = cb200_read_rgba8(self->device, image, width, height, stride, x, y);
        if (rc != CB200_OK) {
            fprintf(stderr, "canvas_b200: cb200_read_rgba8 failed (%d): %s\n", rc,
                    cb200_last_error());
            abort();
        }
    }
}

void canvas::put_image_data(unsigned char const *image, int width, int height, int stride,
                            int x, int y)
{
    if (!image) return;
    self->flush();                                       // ordered after earlier draws
    if (self->tap.write_rgba8)
        self->tap.write_rgba8(self->tap.user, image, width, height, stride, x, y);
    else if (self->device) {
        int rc = cb200_write_rgba8(self->device, image, width, height, stride, x, y);
        if (rc != CB200_OK) {
            fprintf(stderr, "canvas_b200: cb200_write_rgba8 failed (%d): %s\n", rc,
                    cb200_last_error());
            abort();
        }
    }
}

// ------------------------------------------------------------ state stack ----

void canvas::save()
{
    drawing_state st;
    st.op = global_composite_operation;
    st.shadow_offset_x = shadow_offset_x; st.shadow_offset_y = shadow_offset_y;
    st.cap = line_cap; st.join = line_join;
    st.dash_offset = line_dash_offset;
    st.align = text_align; st.baseline = text_baseline;
    st.forward = self->forward; st.inverse = self->inverse;
    st.global_alpha = self->global_alpha;
    st.shadow_color = self->shadow_color;
    st.shadow_blur = self->shadow_blur;
    st.line_width = self->line_width; st.miter_limit = self->miter_limit;
    st.dash = self->dash;
    st.fill = self->fill; st.stroke = self->stroke;
    st.mask = self->mask;                                // masks are immutable device slots
    st.face = self->face;
    self->saves.push_back(st);
}

void canvas::restore()
{
    if (self->saves.empty()) return;
    drawing_state &st = self->saves.back();
    global_composite_operation = st.op;
    shadow_offset_x = st.shadow_offset_x; shadow_offset_y = st.shadow_offset_y;
    line_cap = st.cap; line_join = st.join;
    line_dash_offset = st.dash_offset;
    text_align = st.align; text_baseline = st.baseline;
    self->forward = st.forward; self->inverse = st.inverse;
    self->global_alpha = st.global_alpha;
    self->shadow_color = st.shadow_color;
    self->shadow_blur = st.shadow_blur;
    self->line_width = st.line_width; self->miter_limit = st.miter_limit;
    self->dash.swap(st.dash);
    self->fill = st.fill; self->stroke = st.stroke;
    self->fill.serial = ++self->serial_counter;
    self->stroke.serial = ++self->serial_counter;
    self->mask = st.mask;
    self->face = st.face;
    self->saves.pop_back();
}

}  // namespace canvas_ity
