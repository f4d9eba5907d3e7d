// front_state.hpp -- private state of the recording front end (host C++).
//
// Mirrors what the reference keeps in `class canvas` private members
// (src/canvas_ity.hpp:1150-1172): transform pair, alpha, shadow, line style,
// three brushes, the cubic path, the clip mask (here: a slot id on the device),
// the font tables and the save stack.  What is new is the frame builder: draw
// calls append cb200_draw records plus their pooled data, and flush() hands the
// pools to the back end in one cb200_submit().
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/canvas_b200.h"
#include "../../../include/canvas_ity.hpp"
#include "../geom.cuh"

namespace canvas_ity {

using cb200::affine;
using cb200::vec2;

struct color4 { float r, g, b, a; };

struct brush_state {
    uint32_t type = CB200_BRUSH_COLOR;
    std::vector<color4> colors;      // solid: 1 premultiplied; gradient: stops (straight, linear)
    std::vector<float> stops;
    vec2 start = {0, 0}, end = {0, 0};
    float start_radius = 0, end_radius = 0;
    int width = 0, height = 0;
    uint32_t repetition = 0;
    std::vector<uint8_t> texels;     // pattern: tightly packed straight sRGB RGBA8
    uint64_t serial = 0;             // bumped on every mutation (frame-level dedup)
};

struct glyph_cache;                  // parsed outlines of one font (canvas_front.cpp), shared by content

struct font_state {
    std::vector<uint8_t> data;       // table directory + the 8 tables we use
    int cmap = 0, glyf = 0, head = 0, hhea = 0, hmtx = 0, loca = 0, maxp = 0, os_2 = 0;
    float scale = 0;
    std::shared_ptr<glyph_cache> cache;
};

struct path_state {
    std::vector<vec2> points;                    // device space
    struct sub { uint32_t count; bool closed; };
    std::vector<sub> subs;
};

// Everything save()/restore() snapshot (reference :3410-3465).
struct drawing_state {
    composite_operation op;
    float shadow_offset_x, shadow_offset_y;
    cap_style cap; join_style join;
    float dash_offset;
    align_style align; baseline_style baseline;
    affine forward, inverse;
    float global_alpha;
    color4 shadow_color;
    float shadow_blur, line_width, miter_limit;
    std::vector<float> dash;
    brush_state fill, stroke;
    uint32_t mask;
    font_state face;
};

// Test/bench hook: when installed, frames and pixel reads go to these
// callbacks instead of the CUDA back end (no device is touched).  Only the
// parity tests use it, to hand the very same lowered frame to the oracle.
struct frame_tap {
    void *user = nullptr;
    void (*frame)(void *user, const cb200_frame *frame) = nullptr;
    void (*read_rgba8)(void *user, uint8_t *dst, int w, int h, int stride, int x, int y) = nullptr;
    void (*write_rgba8)(void *user, const uint8_t *src, int w, int h, int stride, int x, int y) = nullptr;
};

struct canvas::host_state {
    int width = 0, height = 0;
    affine forward, inverse;
    float global_alpha = 1.0f;
    color4 shadow_color = {0, 0, 0, 0};
    float shadow_blur = 0, line_width = 1.0f, miter_limit = 10.0f;
    std::vector<float> dash;
    brush_state fill, stroke, image;
    path_state path;
    uint32_t mask = 0;               // current clip-mask slot (0 = whole canvas)
    uint32_t next_mask = 1;
    uint32_t masks_at_last_keep = 1;
    font_state face;
    std::vector<drawing_state> saves;
    uint64_t serial_counter = 1;

    // frame under construction
    std::vector<cb200_draw> draws;
    std::vector<cb200_subpath> subpaths;
    std::vector<float> points;
    std::vector<cb200_brush> brushes;
    std::vector<float> colors, stops, dashes;
    std::vector<cb200_image> images;
    std::vector<uint8_t> texels;
    std::vector<std::shared_ptr<glyph_cache> > frame_atlases;   // fonts the queued text draws refer to
    std::vector<cb200_glyph_inst> glyphs;
    uint32_t n_glyph_points = 0;
    bool instanced_text = true;      // false: lower glyph outlines on the host (A/B parity tests)
    uint64_t cached_brush_serial[3] = {0, 0, 0};
    uint32_t cached_brush_index[3] = {0, 0, 0};
    size_t max_queued_draws = 1u << 16;

    cb200_canvas *device = nullptr;
    frame_tap tap;
    uint64_t frames_flushed = 0;

    void flush();
    void reset_frame();
};

// Host helpers shared with the bindings.
void flattened_path_edges(const canvas::host_state *self, std::vector<float> &edges);
color4 srgb_to_premultiplied_linear(float r, float g, float b, float a);

}  // namespace canvas_ity
