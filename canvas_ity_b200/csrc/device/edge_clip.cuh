// edge_clip.cuh -- clipping and scan conversion of ONE outline edge, host + device.
//
// The reference clips whole polygons against the padded canvas (Sutherland-Hodgman, lines_to_runs
// hpp:2208-2229: left, top, right, bottom), clamps (hpp:2231-2238) and scan-converts every edge of the
// result (add_runs hpp:2109-2170).  A polygon clip is sequential along the loop; here every edge is
// clipped on its own, by the same four stages in the same order with the same arithmetic:
//
//   * clip_edge: the part of an edge that survives all four stages has bit-for-bit the end points the
//     polygon clip gives it (every stage lerps `from` -> `to` of the edge as the previous stage left
//     it, exactly what the polygon clip does with that edge);
//   * what the polygon clip puts in place of the parts cut away is a chain of segments along the
//     boundary.  Along the top and bottom they are horizontal and leave no runs (hpp:2113-2114).  Along
//     the left / right side they carry area: an excursion beside the canvas is replaced by ONE vertical
//     segment from the exit point to the next entry point.  Per edge, the part cut away by the left or
//     right stage is PROJECTED onto that side (x = 0 or x = w, y clamped) instead: the projections of an
//     excursion's edges add up to the same signed length in every scanline;
//   * edge_walk / row_walk / walk_row_runs: add_runs restated per (edge, scanline) -- the reference's
//     DDA evaluates row and column positions directly from the edge's start, so any scanline can start
//     cold and reproduces the reference's deltas bit for bit;
//   * shadow_box_walk: render_shadow (hpp:2409-2419) takes its working rectangle from the bounding box of
//     the polygon clip's runs; projected pieces do not reproduce the boundary segments' extent, so the
//     segments are rebuilt from the loop's crossings in one ordered walk (see there).
//
// Must be compiled without FMA contraction (-fmad=false / -ffp-contract=off).
#pragma once

#include "../geom.cuh"

#include <cuda_runtime.h>

namespace cb200 {

enum { BOX_LEFT = 0, BOX_RIGHT = 1, BOX_TOP = 2, BOX_OUTWARD = 4 };   // box_event::kind = line | outward flag

struct box_event {
    float t;         // parameter along the (original) edge, for ordering only
    float v;         // y of a side crossing, x of a top crossing
    int kind;
};

struct clipped_edge {
    int n_pieces;                 // <= 3: left projection, inside part, right projection
    float4 piece[3];              // (x0, y0, x1, y1), in the edge's direction
    int projected[3];
    int n_events;                 // crossings of x = 0, y = 0, x = w as the polygon clip sees them, in edge order
    box_event ev[3];
};

// One clip stage on the edge a -> b: `da`, `db` are the signed distances of its ends from the clip
// line (>= 0 is kept, hpp:2223).  Returns 0 both ends kept, 1 both cut away, 2 crossing: then `at` is the
// clip's own lerp (hpp:2220-2222) -- or the kept end itself when that lies exactly on the line and the
// clip inserts nothing -- and `t` its parameter on a -> b.
CB_HD int clip_stage(vec2 a, vec2 b, float da, float db, float &t, vec2 &at, bool &a_kept)
{
    const bool in_a = da >= 0.0f, in_b = db >= 0.0f;
    a_kept = in_a;
    if (in_a && in_b) return 0;
    if (!in_a && !in_b) return 1;
    if (da * db < 0.0f) { t = da / (da - db); at = mix(a, b, t); }
    else { t = in_a ? 0.0f : 1.0f; at = in_a ? a : b; }
    return 2;
}

CB_HD float clamp_to(float v, float hi) { return fminf(fmaxf(v, 0.0f), hi); }

// crossings are found stage by stage; order them along the edge (at most three)
CB_HD void sort_events(clipped_edge &out)
{
    if (out.n_events >= 2 && out.ev[1].t < out.ev[0].t) { box_event s = out.ev[0]; out.ev[0] = out.ev[1]; out.ev[1] = s; }
    if (out.n_events == 3) {
        if (out.ev[2].t < out.ev[1].t) { box_event s = out.ev[1]; out.ev[1] = out.ev[2]; out.ev[2] = s; }
        if (out.ev[1].t < out.ev[0].t) { box_event s = out.ev[0]; out.ev[0] = out.ev[1]; out.ev[1] = s; }
    }
}

CB_HD void clip_edge(vec2 a, vec2 b, float w, float h, clipped_edge &out)
{
    out.n_pieces = 0;
    out.n_events = 0;
    float ta = 0.0f, tb = 1.0f;                       // the current part's range on the original edge
    float t = 0.0f;
    vec2 at = a;
    bool a_kept = true;
    auto project = [&](float col, float y0, float y1) {
        const int k = out.n_pieces++;
        out.piece[k].x = col; out.piece[k].y = clamp_to(y0, h); out.piece[k].z = col; out.piece[k].w = clamp_to(y1, h);
        out.projected[k] = 1;
    };
    auto event = [&](int line, float v) {
        const int k = out.n_events++;
        out.ev[k].t = ta + t * (tb - ta); out.ev[k].v = v; out.ev[k].kind = line | (a_kept ? BOX_OUTWARD : 0);
    };
    auto shrink = [&]() {
        const float tx = ta + t * (tb - ta);
        if (a_kept) { b = at; tb = tx; } else { a = at; ta = tx; }
    };
    bool alive = true;
    // left
    int r = clip_stage(a, b, a.x, b.x, t, at, a_kept);
    if (r == 1) { project(0.0f, a.y, b.y); alive = false; }
    else if (r == 2) {
        event(BOX_LEFT, at.y);
        if (a_kept) project(0.0f, at.y, b.y); else project(0.0f, a.y, at.y);
        shrink();
    }
    // top: what is cut away becomes a horizontal boundary segment, no area
    if (alive) {
        r = clip_stage(a, b, a.y, b.y, t, at, a_kept);
        if (r == 1) alive = false;
        else if (r == 2) { event(BOX_TOP, at.x); shrink(); }
    }
    // right
    if (alive) {
        r = clip_stage(a, b, w - a.x, w - b.x, t, at, a_kept);
        if (r == 1) { project(w, a.y, b.y); alive = false; }
        else if (r == 2) {
            event(BOX_RIGHT, at.y);
            if (a_kept) project(w, at.y, b.y); else project(w, a.y, at.y);
            shrink();
        }
    }
    // bottom
    if (alive) {
        r = clip_stage(a, b, h - a.y, h - b.y, t, at, a_kept);
        if (r == 1) alive = false;
        else if (r == 2) shrink();
    }
    if (alive) {
        const int k = out.n_pieces++;
        out.piece[k].x = clamp_to(a.x, w); out.piece[k].y = clamp_to(a.y, h);
        out.piece[k].z = clamp_to(b.x, w); out.piece[k].w = clamp_to(b.y, h);
        out.projected[k] = 0;
    }
    sort_events(out);
}

// ------------------------------------------------------------ add_runs, per scanline ----

// add_runs drops an edge flatter than 2e-5 (hpp:2113-2114).  A projected piece is exempt: the projections of an
// excursion form a closed chain with the inside pieces only if every link is kept -- the reference has ONE boundary
// segment there (or, for a loop wholly beside the canvas, nothing at all), so a dropped link would leave a coverage
// residue of up to 2e-5 per loop that the reference does not have (round-1 fuzz seed 672: ten glyph outlines left
// of the canvas, each with one 1.5e-5 link, add up to more than the 1/8160 paint threshold along a whole row).
CB_HD bool piece_has_runs(float4 pc, bool projected)
{
    return projected ? pc.w != pc.y : !(fabsf(pc.w - pc.y) < 2.0e-5f);
}

struct edge_walk {
    vec2 from, to;
    float sign, ystep, dxdy, dydx, fx0, fy0;
    bool vertical, down;
    int rows;                  // scanlines the reference loop would visit
};

CB_HD edge_walk edge_setup(float4 pc)
{
    edge_walk e;
    vec2 a = v2(pc.x, pc.y), b = v2(pc.z, pc.w);
    e.sign = b.y > a.y ? 1.0f : -1.0f;
    if (a.x > b.x) { vec2 t = a; a = b; b = t; }          // always walk left to right
    e.from = a; e.to = b;
    e.down = b.y > a.y;
    e.ystep = e.down ? 1.0f : -1.0f;
    e.dxdy = (b.x - a.x) / (b.y - a.y);
    e.dydx = (b.y - a.y) / (b.x - a.x);
    e.vertical = b.x - a.x < 2.0e-5f;
    e.fx0 = floorf(a.x);
    e.fy0 = floorf(a.y);
    e.rows = e.down ? int(ceilf(b.y) - e.fy0) : int(e.fy0 - floorf(b.y)) + 1;
    return e;
}

struct row_walk {
    vec2 now, stop;            // entry / exit of the edge in this scanline
    float px, py;              // first pixel touched
    int inner;                 // pixels crossed before the last one
};

CB_HD float edge_x_at(const edge_walk &e, float y) { return (y - e.from.y) * e.dxdy + e.from.x; }
CB_HD float edge_y_at(const edge_walk &e, float x) { return (x - e.from.x) * e.dydx + e.from.y; }

CB_HD row_walk row_setup(const edge_walk &e, int r)
{
    row_walk w;
    float fr = float(r);
    if (e.down) {
        w.py = e.fy0 + fr;
        float y_in = e.fy0 + fr, y_out = e.fy0 + fr + 1.0f;
        w.now = r == 0 ? e.from : v2(edge_x_at(e, y_in), y_in);
        w.stop = e.to.y < y_out ? e.to : v2(edge_x_at(e, y_out), y_out);
    } else {
        w.py = e.fy0 - fr;
        float y_in = e.fy0 - fr + 1.0f, y_out = e.fy0 - fr;
        w.now = r == 0 ? e.from : v2(edge_x_at(e, y_in), y_in);
        w.stop = e.to.y > y_out ? e.to : v2(edge_x_at(e, y_out), y_out);
    }
    w.px = (r == 0 || e.vertical) ? e.fx0 : fmaxf(e.fx0, ceilf(w.now.x) - 1.0f);
    float crossed = e.vertical ? 0.0f : ceilf(w.stop.x) - 1.0f - w.px;
    w.inner = crossed > 0.0f ? int(crossed) : 0;
    return w;
}

// The runs of one scanline of one edge, left to right: sink.put(px, delta), w.inner + 2 of them.
template <class Sink>
CB_HD void walk_row_runs(const edge_walk &e, const row_walk &w, Sink &sink)
{
    vec2 cur = w.now;
    float px = w.px, carry_area = 0.0f;
    for (int c = 0; c < w.inner; ++c) {
        float gx = px + 1.0f;
        vec2 nx = v2(gx, edge_y_at(e, gx));
        float strip = clamp01((nx.y - cur.y) * e.ystep);
        float mid = (nx.x + cur.x) * 0.5f;
        float area = (mid - px) * strip;
        sink.put(px, (carry_area + strip - area) * e.sign);
        carry_area = area;
        cur = nx;
        px = gx;
    }
    float strip = clamp01((w.stop.y - cur.y) * e.ystep);
    float mid = (w.stop.x + cur.x) * 0.5f;
    float area = (mid - px) * strip;
    sink.put(px, (carry_area + strip - area) * e.sign);
    sink.put(px + 1.0f, area * e.sign);
}

// ------------------------------------------------------------ shadow working rectangle ----
//
// What the reference's polygon clip adds to the run bounding box of a shadow beyond the runs of the
// inside pieces: the vertical boundary segments.  Rebuilt from the loop's crossings (clip_edge's events)
// in one ordered walk over the loop's edges:
//
//   * crossings of one clip line alternate outward / inward along a closed loop; every outward crossing
//     pairs with the next crossing of that line (cyclically) and the pair is the boundary segment;
//   * the clip goes left, top, right, bottom.  By the time the top is clipped everything left of x = 0 is
//     gone (clip_edge finds top crossings on the left-clipped edge), but a left boundary segment whose
//     ends lie either side of y = 0 crosses it at x = 0.  By the time the right side is clipped everything
//     above y = 0 is gone, and a top boundary segment (between a crossing that leaves through y = 0 and
//     the next return) that straddles x = w is cut at (w, 0): a crossing of the right side at y = 0;
//   * the top / bottom clip and the final clamp confine a segment to [0, h]; a segment wholly above or
//     wholly below becomes horizontal and disappears;
//   * a vertical segment from y_lo to y_hi at column c leaves non-zero runs in rows
//     floor(y_lo) .. ceil(y_hi) - 1 of column c (add_runs with area = 0).
//
// The kernel (raster.cu, k_shadow_boxes) gives one warp to a loop: lanes clip 32 edges at a time, then
// the few lanes that found crossings feed them, in edge order, to this walk (state replicated in every
// lane).  cb200_debug_shadow_box (backend.cu) runs the same code serially on the host for the CPU test
// against the reference's clip (tests/test_random_scenes.py).

struct box_line {
    int have, prev_outward, have_first;
    float prev_v, first_v;
};

struct shadow_box_walk {
    box_line left, right, top;
    float w, h;                       // padded canvas
    int lx, hx, ly, hy;               // what the boundary segments add to the run bounding box

    CB_HD void init(float width, float height)
    {
        left.have = left.prev_outward = left.have_first = 0; left.prev_v = left.first_v = 0.0f;
        right = left; top = left;
        w = width; h = height;
        lx = ly = 0x7fffffff; hx = hy = -1;
    }

    // the boundary segment between side crossings at y1 and y2, column `col`
    CB_HD void segment(float y1, float y2, int col)
    {
        if ((y1 < 0.0f && y2 < 0.0f) || (y1 > h && y2 > h)) return;
        const float c1 = clamp_to(y1, h), c2 = clamp_to(y2, h);
        const float lo = fminf(c1, c2), hi = fmaxf(c1, c2);
        if (!(hi - lo >= 2.0e-5f)) return;                       // add_runs drops it, hpp:2113-2114
        const int r_lo = int(floorf(lo)), r_hi = int(ceilf(hi)) - 1;
        lx = col < lx ? col : lx; hx = col > hx ? col : hx;
        ly = r_lo < ly ? r_lo : ly; hy = r_hi > hy ? r_hi : hy;
    }

    CB_HD void feed_right(float y, bool outward)
    {
        if (right.have && right.prev_outward) segment(right.prev_v, y, int(w));
        right.have = 1; right.prev_outward = outward ? 1 : 0; right.prev_v = y;
        if (!right.have_first) { right.have_first = 1; right.first_v = y; }
    }

    // the top boundary segment between crossings of y = 0 at xa (leaving) and xb (returning)
    CB_HD void top_segment(float xa, float xb)
    {
        if ((w - xa) * (w - xb) < 0.0f) feed_right(0.0f, xa < w);
    }

    CB_HD void feed_top(float x, bool leaves)
    {
        if (top.have && top.prev_outward) top_segment(top.prev_v, x);
        top.have = 1; top.prev_outward = leaves ? 1 : 0; top.prev_v = x;
        if (!top.have_first) { top.have_first = 1; top.first_v = x; }
    }

    // a left boundary segment from y1 (exit) to y2 (entry): its runs, and its own crossing of y = 0
    CB_HD void left_segment(float y1, float y2)
    {
        segment(y1, y2, 0);
        if ((y1 >= 0.0f) != (y2 >= 0.0f)) feed_top(0.0f, y1 >= 0.0f);
    }

    CB_HD void consume(int kind, float v)
    {
        const bool outward = (kind & BOX_OUTWARD) != 0;
        const int line = kind & 3;
        if (line == BOX_LEFT) {
            if (left.have && left.prev_outward) left_segment(left.prev_v, v);
            left.have = 1; left.prev_outward = outward ? 1 : 0; left.prev_v = v;
            if (!left.have_first) { left.have_first = 1; left.first_v = v; }
        } else if (line == BOX_RIGHT) feed_right(v, outward);
        else feed_top(v, outward);
    }

    // Close the loop: the last outward crossing of every line pairs with that line's first crossing.
    // What a wrap-around pair feeds onward arrives last although it belongs to the loop's start; the
    // sequences are cyclic and nothing that is kept can lie in between (the loop is outside there), so
    // the pairs come out the same.
    CB_HD void finish()
    {
        if (left.have && left.prev_outward) left_segment(left.prev_v, left.first_v);
        if (top.have && top.prev_outward) top_segment(top.prev_v, top.first_v);
        if (right.have && right.prev_outward) segment(right.prev_v, right.first_v, int(w));
    }
};

}  // namespace cb200
