// geometry.cu -- K1 flatten, K2 dash, K3 stroke expansion for ALL draws of a
// frame at once (sm_100a).  Compiled with -fmad=false: every decision below is
// a float comparison that must agree with the reference's unfused arithmetic
// (SURVEY 7.4).
//
//   K1  one thread per flatten unit (subpath start point or one cubic):
//       count pass -> in-kernel partial scan -> emit pass, so output order is the
//       reference's point order (add_bezier / add_tessellation, hpp:1331-1524).
//   K2  one thread per dashed source subpath: the dash phase walk is a serial
//       recurrence in the reference (hpp:1858-1934) and is kept serial per subpath;
//       subpaths run in parallel.
//   K3  one thread per stroke unit (one join or one cap of one half stroke), same
//       count/emit structure (add_half_stroke / stroke_lines, hpp:1949-2100); the
//       reference's serial recurrence is only replayed (by one warp) for polylines
//       that contain a segment shorter than 1e-4.  Round joins and caps call the K1
//       device routine for their arcs.
#include "frame.cuh"

namespace cb200 {

namespace {

__device__ __forceinline__ vec2 ld(const float2 *p, uint32_t i) { float2 v = p[i]; return v2(v.x, v.y); }

// ------------------------------------------------------------------- K0 ----
// Glyph instances -> device-space cubics (device-side text).  One warp per instance,
// one lane per piece of the cached outline.  Per piece this is the arithmetic of the host lowering
// (canvas_front.cpp lower_glyph, which restates add_glyph hpp:1533-1696): transform the font-unit
// points it refers to, take the implied midpoints, degree-elevate the quadratic -- same products
// in the same order without FMA, so the points equal the uploaded ones bit for bit.
__global__ void __launch_bounds__(kBlock) k_glyph_instances(device_frame f)
{
    grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (kBlock / 32);
    for (uint32_t g = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); g < f.n_glyph_insts; g += warps) {
        const glyph_inst_rec gi = f.glyph_insts[g];
        const atlas_dev at = f.atlas_table[gi.atlas];
        const cb200_glyph_outline o = at.outlines[gi.outline];
        const float2 *src = at.points + o.first_point;
        float2 *dst = f.in_points + gi.first_point;
        auto point = [&](uint32_t k) { const float2 u = src[k]; return apply(gi.m, v2(u.x, u.y)); };
        auto end_point = [&](uint32_t a, uint32_t b) { return a == b ? point(a) : mix(point(a), point(b), 0.5f); };
        for (uint32_t k = lane; k < o.n_segs; k += 32) {
            const cb200_glyph_seg sg = at.segs[o.first_seg + k];
            const vec2 from = end_point(sg.from_a, sg.from_b), to = end_point(sg.to_a, sg.to_b);
            vec2 c1 = from, c2 = to;                                   // straight: (from, to, to)
            if (!(sg.flags & CB200_SEG_LINE)) {
                const vec2 c = point(sg.ctrl);
                c1 = mix(from, c, 2.0f / 3.0f);
                c2 = mix(to, c, 2.0f / 3.0f);
            }
            if (sg.flags & CB200_SEG_FIRST) dst[sg.out - 1] = make_float2(from.x, from.y);
            dst[sg.out] = make_float2(c1.x, c1.y);
            dst[sg.out + 1] = make_float2(c2.x, c2.y);
            dst[sg.out + 2] = make_float2(to.x, to.y);
        }
    }
}

// ------------------------------------------------------------------- K1 ----

struct point_sink {
    float2 *out; uint32_t *loop_of; uint32_t at, loop;
    __device__ __forceinline__ void put(vec2 p) { out[at] = make_float2(p.x, p.y); loop_of[at] = loop; ++at; }
};

constexpr int kWarps = kBlock / 32;
constexpr int kNodeWords = 10;          // p0 c1 c2 p3 (8 floats) + budget + state

// One flatten unit by one WARP.  The reference recursion (add_tessellation) is a
// depth-first walk whose subtrees are independent, so the warp first expands the
// tree breadth-first -- the monotone pieces of add_bezier are the initial nodes,
// every level halves the nodes that fail the flatness test, a warp scan keeps the
// node list in curve order -- until up to 32 nodes exist; then every lane finishes
// its own subtree depth-first and a second scan orders the output.  Same halving
// arithmetic and the same tests as the serial routine in geom.cuh, hence the same
// points in the same order.  `nodes`: kNodeWords * 32 words of warp-private smem.
template <bool EMIT>
__device__ uint32_t warp_flatten(const device_frame &f, uint32_t u, uint32_t out_base, float *nodes)
{
    const int lane = threadIdx.x & 31;
    unit_rec un = f.units[u];
    subpath_rec sp = f.subpaths[un.subpath];
    if (un.index == 0) {                                  // the subpath's start point
        if (EMIT && lane == 0) {
            point_sink ps = { f.pts, f.pt_loop, out_base, un.subpath };
            ps.put(ld(f.in_points, sp.first_point));
        }
        return 1;
    }
    const uint32_t at = sp.first_point + 3 * (un.index - 1);
    const vec2 P0 = ld(f.in_points, at), C1 = ld(f.in_points, at + 1), C2 = ld(f.in_points, at + 2),
               P3 = ld(f.in_points, at + 3);
    const float angular = f.draws[sp.draw].angular;
    {
        vec2 e1 = C1 - P0, e3 = P3 - C2;
        if (dot(e1, e1) == 0.0f && dot(e3, e3) == 0.0f) { // a straight segment
            if (EMIT && lane == 0) {
                point_sink ps = { f.pts, f.pt_loop, out_base, un.subpath };
                ps.put(P3);
            }
            return 1;
        }
    }
    // initial nodes: the kept monotone pieces, piece m on lane m
    vec2 p0 = v2(0, 0), c1 = p0, c2 = p0, p3 = p0;
    int budget = 0, state = 0;                            // 0 empty, 1 open, 2 flat leaf
    {
        float cut[7];
        int n = cubic_cuts(P0, C1, C2, P3, cut), m = 0;
        vec2 from = P0;
        for (int i = 0; i + 1 < n; ++i) {
            if (!cut_is_kept(cut[i], cut[i + 1])) continue;
            vec2 k1, k2, to;
            cubic_piece(P0, C1, C2, P3, cut[i], cut[i + 1], k1, k2, to);
            if (lane == m) { p0 = from; c1 = k1; c2 = k2; p3 = to; budget = 20; state = 1; }
            from = to;
            ++m;
        }
    }
    for (;;) {                                            // breadth-first expansion
        float q1, q2, q3;
        if (state == 1 && (piece_is_flat(p0, c1, c2, p3, angular, q1, q2, q3) || budget == 0)) state = 2;
        uint32_t width = state == 0 ? 0u : (state == 1 ? 2u : 1u);
        uint32_t incl = warp_inclusive_scan(width);
        uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (!__any_sync(0xffffffffu, state == 1) || total > 32) break;
        uint32_t pos = incl - width;
        if (state == 2) {
            float *d = nodes + pos * kNodeWords;
            d[0] = p0.x; d[1] = p0.y; d[2] = c1.x; d[3] = c1.y; d[4] = c2.x; d[5] = c2.y; d[6] = p3.x; d[7] = p3.y;
            d[8] = __int_as_float(budget); d[9] = __int_as_float(2);
        } else if (state == 1) {
            vec2 l1 = mix(p0, c1, 0.5f), mid = mix(c1, c2, 0.5f), r2 = mix(c2, p3, 0.5f);
            vec2 l2 = mix(l1, mid, 0.5f), r1 = mix(mid, r2, 0.5f);
            vec2 split = mix(l2, r1, 0.5f);
            float *d = nodes + pos * kNodeWords;
            d[0] = p0.x; d[1] = p0.y; d[2] = l1.x; d[3] = l1.y; d[4] = l2.x; d[5] = l2.y; d[6] = split.x; d[7] = split.y;
            d[8] = __int_as_float(budget - 1); d[9] = __int_as_float(1);
            d += kNodeWords;
            d[0] = split.x; d[1] = split.y; d[2] = r1.x; d[3] = r1.y; d[4] = r2.x; d[5] = r2.y; d[6] = p3.x; d[7] = p3.y;
            d[8] = __int_as_float(budget - 1); d[9] = __int_as_float(1);
        }
        __syncwarp();
        if (uint32_t(lane) < total) {
            const float *d = nodes + lane * kNodeWords;
            p0 = v2(d[0], d[1]); c1 = v2(d[2], d[3]); c2 = v2(d[4], d[5]); p3 = v2(d[6], d[7]);
            budget = __float_as_int(d[8]); state = __float_as_int(d[9]);
        } else state = 0;
        __syncwarp();
    }
    // depth-first finish of every lane's subtree; count first, then write in order
    uint32_t mine = 0;
    if (state == 2) {
        float q1, q2, q3;
        piece_is_flat(p0, c1, c2, p3, angular, q1, q2, q3);
        count_sink cs = { 0 };
        emit_flat_piece(c1, c2, p3, angular, q1, q2, q3, cs);
        mine = uint32_t(cs.n);
    } else if (state == 1) {
        count_sink cs = { 0 };
        subdivide_piece(p0, c1, c2, p3, angular, cs, budget);
        mine = uint32_t(cs.n);
    }
    uint32_t incl = warp_inclusive_scan(mine);
    if (EMIT && mine) {
        point_sink ps = { f.pts, f.pt_loop, out_base + incl - mine, un.subpath };
        if (state == 2) {
            float q1, q2, q3;
            piece_is_flat(p0, c1, c2, p3, angular, q1, q2, q3);
            emit_flat_piece(c1, c2, p3, angular, q1, q2, q3, ps);
        } else
            subdivide_piece(p0, c1, c2, p3, angular, ps, budget);
    }
    return __shfl_sync(0xffffffffu, incl, 31);
}

__global__ void __launch_bounds__(kBlock) k_flatten_count(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    __shared__ uint32_t block_total;
    __shared__ float nodes[kWarps][kNodeWords * 32];
    uint32_t n = f.hdr->n_units, begin, end;
    warp_slice(n, begin, end);
    if (threadIdx.x == 0) block_total = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t u = begin + uint32_t(warp); u < end; u += kWarps) {
        uint32_t c = warp_flatten<false>(f, u, 0, nodes[warp]);
        if (lane == 0) { f.unit_count[u] = c; atomicAdd(&block_total, c); }
    }
    __syncthreads();
    if (threadIdx.x == 0) f.partials[blockIdx.x] = block_total;
    finish_partials(f.partials, &f.hdr->tickets[0], &f.hdr->n_line_points, sm);
}

__global__ void __launch_bounds__(kBlock) k_flatten_emit(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    __shared__ float nodes[kWarps][kNodeWords * 32];
    uint32_t n = f.hdr->n_units, begin, end;
    warp_slice(n, begin, end);
    if (f.hdr->n_line_points > f.cap_pts) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&f.hdr->overflow, OVF_POINTS);
        return;
    }
    uint32_t carry = f.partials[blockIdx.x];
    for (uint32_t chunk = begin; chunk < end; chunk += kBlock) {
        uint32_t u = chunk + threadIdx.x;
        uint32_t c = u < end ? f.unit_count[u] : 0, total;
        uint32_t ex = block_exclusive_scan(c, sm, total);
        if (u < end) f.unit_offset[u] = carry + ex;
        if (u < end && u == n - 1) f.unit_offset[n] = carry + ex + c;
        carry += total;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    for (uint32_t u = begin + uint32_t(warp); u < end; u += kWarps)
        warp_flatten<true>(f, u, f.unit_offset[u], nodes[warp]);
}

// loops[s] = point span of subpath s in the K1 output
__global__ void k_subpath_loops(device_frame f)
{
    grid_dependency_wait();
    uint32_t n = f.hdr->n_subpaths;
    if (f.hdr->overflow) return;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        subpath_rec sp = f.subpaths[s];
        uint32_t a = f.unit_offset[sp.first_unit], b = f.unit_offset[sp.first_unit + 1 + sp.n_cubics];
        loop_span l = { a, b - a };
        f.loops[s] = l;
    }
}

// ------------------------------------------------------------------- K2 ----

// Serial dash walk of one polyline.  Sink: pt(vec2) appends a point to the dash
// under construction, cut() ends it.  Returns through `st` what the closed-path
// fix-up needs.
struct dash_stats { uint32_t points, dashes, tail; bool single_closed; };

template <class Sink>
__device__ void dash_walk(const device_frame &f, const draw_rec &d, loop_span src, bool closed,
                          Sink &sink, dash_stats &st)
{
    const float *pat = f.dashes + d.first_dash;
    const uint32_t np = d.n_dash;
    float total = 0.0f;
    for (uint32_t i = 0; i < np; ++i) total += pat[i];
    float phase = fmodf(d.dash_offset, total);
    if (phase < 0.0f) phase += total;
    uint32_t seg0 = 0;
    while (phase >= pat[seg0]) {
        phase -= pat[seg0];
        seg0 = seg0 + 1 < np ? seg0 + 1 : 0;
    }
    uint32_t seg = seg0;
    bool on = (seg0 & 1u) == 0;
    const bool began_on = on;
    float until = pat[seg0] - phase;
    uint32_t dashes = 0, in_dash = 0, points = 0;
    uint32_t i = src.first;
    for (; i + 1 < src.first + src.count; ++i) {
        vec2 a = ld(f.pts, i), b = ld(f.pts, i + 1);
        if (on) { sink.pt(a); ++in_dash; ++points; }
        float len = vlen(apply(d.inverse, b) - apply(d.inverse, a));   // user-space length
        while (until < len) {
            sink.pt(mix(a, b, until / len)); ++in_dash; ++points;
            if (on) { sink.cut(in_dash); ++dashes; in_dash = 0; }
            seg = seg + 1 < np ? seg + 1 : 0;
            on = !on;
            until += pat[seg];
        }
        until -= len;
    }
    st.tail = 0;
    st.single_closed = false;
    if (on) {
        sink.pt(ld(f.pts, i)); ++in_dash; ++points;
        sink.cut(in_dash); ++dashes;
        if (closed && began_on) {
            if (dashes == 1) st.single_closed = true;      // one dash spans the whole loop
            else { st.tail = in_dash; --dashes; }          // last dash joins the first
        }
    }
    st.points = points;
    st.dashes = dashes;
}

struct dash_count_sink {
    __device__ __forceinline__ void pt(vec2) {}
    __device__ __forceinline__ void cut(uint32_t) {}
};

__global__ void __launch_bounds__(kBlock) k_dash_count(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    uint32_t n = f.n_dash_items, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t first = begin + threadIdx.x * ipt, sum_p = 0, sum_s = 0;
    bool bad = f.hdr->overflow != 0;
    for (uint32_t k = 0; k < ipt && !bad; ++k) {
        uint32_t it = first + k;
        if (it >= end) break;
        uint32_t s = f.dash_items[it].subpath;
        subpath_rec sp = f.subpaths[s];
        dash_count_sink cs;
        dash_stats st;
        dash_walk(f, f.draws[sp.draw], f.loops[s], sp.closed != 0, cs, st);
        f.dash_pts_count[it] = st.points;
        f.dash_sub_count[it] = st.dashes;
        f.dash_tail[it] = st.tail | (st.single_closed ? 0x80000000u : 0u);
        sum_p += st.points;
        sum_s += st.dashes;
    }
    uint32_t total;
    block_exclusive_scan(sum_p, sm, total);
    if (threadIdx.x == 0) f.partials[kGrid + blockIdx.x] = total;
    block_exclusive_scan(sum_s, sm, total);
    if (threadIdx.x == 0) f.partials[2 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + kGrid, &f.hdr->tickets[1], &f.hdr->n_dash_points, sm);
    finish_partials(f.partials + 2 * kGrid, &f.hdr->tickets[2], &f.hdr->n_dash_subpaths, sm);
}

// Writes points with the closed-path rotation applied on the fly: the last
// `tail` points (the final dash) land in front of the first dash.
struct dash_emit_sink {
    float2 *out; loop_span *loops; stroke_src *sources;
    uint32_t base, total, tail;         // point base / count / rotated tail of this source
    uint32_t k;                         // running point index
    uint32_t loop0, src0, m, kept;      // first loop id / source slot, dash index, dashes kept
    uint32_t draw_closed, dash_start;
    __device__ __forceinline__ void pt(vec2 p)
    {
        uint32_t dst = tail ? (k >= total - tail ? k - (total - tail) : k + tail) : k;
        out[base + dst] = make_float2(p.x, p.y);
        ++k;
    }
    __device__ __forceinline__ void cut(uint32_t count)
    {
        if (m < kept) {
            loop_span l = { base + dash_start + (m ? tail : 0), count + (m == 0 ? tail : 0) };
            loops[loop0 + m] = l;
            stroke_src s = { loop0 + m, draw_closed };
            sources[src0 + m] = s;
        }
        dash_start += count;
        ++m;
    }
};

__global__ void __launch_bounds__(kBlock) k_dash_emit(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    uint32_t n = f.n_dash_items, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    frame_header *h = f.hdr;
    uint32_t pt_base = h->n_line_points;
    if (pt_base + h->n_dash_points > f.cap_pts || h->n_subpaths + h->n_dash_subpaths > f.stroke_loop_base ||
        f.n_static_sources + h->n_dash_subpaths > f.cap_sources) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&h->overflow, OVF_DASH);
        return;
    }
    if (h->overflow) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) h->n_sources = f.n_static_sources + h->n_dash_subpaths;
    uint32_t first = begin + threadIdx.x * ipt, sum_p = 0, sum_s = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) {
        sum_p += f.dash_pts_count[first + k];
        sum_s += f.dash_sub_count[first + k];
    }
    uint32_t total;
    uint32_t at_p = block_exclusive_scan(sum_p, sm, total) + f.partials[kGrid + blockIdx.x];
    uint32_t at_s = block_exclusive_scan(sum_s, sm, total) + f.partials[2 * kGrid + blockIdx.x];
    for (uint32_t k = 0; k < ipt; ++k) {
        uint32_t it = first + k;
        if (it >= end) break;
        dash_item di = f.dash_items[it];
        subpath_rec sp = f.subpaths[di.subpath];
        uint32_t tail_word = f.dash_tail[it];
        dash_emit_sink es;
        es.out = f.pts; es.loops = f.loops; es.sources = f.sources;
        es.base = pt_base + at_p; es.total = f.dash_pts_count[it]; es.tail = tail_word & 0x7fffffffu;
        es.k = 0; es.loop0 = h->n_subpaths + at_s; es.src0 = f.n_static_sources + at_s;
        es.m = 0; es.kept = f.dash_sub_count[it];
        es.draw_closed = sp.draw | (tail_word & 0x80000000u);
        es.dash_start = 0;
        dash_stats st;
        dash_walk(f, f.draws[sp.draw], f.loops[di.subpath], sp.closed != 0, es, st);
        if (di.flags & 1) f.draw_src[sp.draw].x = f.n_static_sources + at_s;
        if (di.flags & 2) f.draw_src[sp.draw].y = f.n_static_sources + at_s + es.kept;   // end (exclusive)
        at_p += es.total;
        at_s += es.kept;
    }
}

// ------------------------------------------------------------------- K3 ----

struct stroke_style {
    float half, miter2; uint32_t cap, join;
    affine fwd, inv;
};

__device__ __forceinline__ stroke_style style_of(const draw_rec &d)
{
    stroke_style st;
    st.half = d.line_width * 0.5f;
    st.miter2 = d.miter_limit * d.miter_limit * st.half * st.half;
    st.cap = d.cap; st.join = d.join;
    st.fwd = d.forward; st.inv = d.inverse;
    return st;
}

// State carried along a half-stroke walk (hpp:1956-1958, 2019-2024): the last
// accepted point (user space), the unit tangent and length of the segment that
// arrived there.
struct walk_state { vec2 pivot, tin; float lin; };

// One step of the walk: the join at `ws.pivot` between the incoming segment and
// pivot->nxt (hpp:1963-2024).  Returns true when a join was emitted.
template <class Sink>
__device__ __forceinline__ bool join_step(walk_state &ws, vec2 nxt, const stroke_style &st, Sink &sink)
{
    const float eps = 1.0e-4f;
    vec2 tin = ws.tin, pivot = ws.pivot;
    vec2 tout = unit(nxt - pivot);
    float lout = vlen(nxt - pivot);
    bool joined = false;
    if (ws.lin != 0.0f && lout >= eps) {
        joined = true;
        vec2 a = pivot + st.half * perp(tin);
        vec2 b = pivot + st.half * perp(tout);
        float turn = dot(perp(tin), tout);
        if (fabsf(turn) < eps) turn = 0.0f;
        vec2 tip = turn == 0.0f ? v2(0.0f, 0.0f) : (st.half / turn) * (tout - tin);
        bool tight = dot(tip, tin) < -ws.lin && dot(tip, tout) > lout;
        bool wrap = turn > 0.0f && tight;            // inner join tighter than the segments
        vec2 jin = tin, jout = tout;
        if (wrap) {
            vec2 t = a; a = b; b = t;
            jin = tout; jout = tin;
            sink.put(apply(st.fwd, b));
            sink.put(apply(st.fwd, pivot));
            sink.put(apply(st.fwd, a));
        }
        if ((turn > 0.0f && !tight) || (turn != 0.0f && st.join == 0 && dot(tip, tip) <= st.miter2))
            sink.put(apply(st.fwd, pivot + tip));
        else if (st.join == 2) {
            float cosine = dot(jin, jout);
            float angle = join_acosf(fminf(fmaxf(cosine, -1.0f), 1.0f));     // the host libm's bits (geom.cuh)
            float k = 4.0f / 3.0f * join_tanf(0.25f * angle);
            sink.put(apply(st.fwd, a));
            flatten_cubic(apply(st.fwd, a), apply(st.fwd, a + (k * st.half) * jin),
                          apply(st.fwd, b - (k * st.half) * jout), apply(st.fwd, b), -1.0f, sink);
        } else {
            sink.put(apply(st.fwd, a));
            sink.put(apply(st.fwd, b));
        }
        if (wrap) {
            sink.put(apply(st.fwd, b));
            sink.put(apply(st.fwd, pivot));
            sink.put(apply(st.fwd, a));
        }
    }
    if (lout >= eps) { ws.tin = tout; ws.lin = lout; ws.pivot = nxt; }
    return joined;
}

// Line cap at the end of an open half (hpp:2029-2057).
template <class Sink>
__device__ __forceinline__ void cap_end(const walk_state &ws, const stroke_style &st, Sink &sink)
{
    vec2 pivot = ws.pivot;
    vec2 ahead = st.half * ws.tin;
    vec2 side = perp(ahead);
    if (st.cap == 0) {                                   // butt
        sink.put(apply(st.fwd, pivot + side));
        sink.put(apply(st.fwd, pivot - side));
    } else if (st.cap == 1) {                            // square
        sink.put(apply(st.fwd, pivot + ahead + side));
        sink.put(apply(st.fwd, pivot + ahead - side));
    } else if (st.cap == 2) {                            // circle: two quarter arcs
        const float k = 0.55228475f;
        sink.put(apply(st.fwd, pivot + side));
        flatten_cubic(apply(st.fwd, pivot + side), apply(st.fwd, pivot + side + k * ahead),
                      apply(st.fwd, pivot + ahead + k * side), apply(st.fwd, pivot + ahead), -1.0f, sink);
        flatten_cubic(apply(st.fwd, pivot + ahead), apply(st.fwd, pivot + ahead - k * side),
                      apply(st.fwd, pivot - side + k * ahead), apply(st.fwd, pivot - side), -1.0f, sink);
    }
}

// One half stroke by one WARP.  The reference walks the polyline serially, but
// the state it carries is a pure function of the two preceding points whenever no
// segment shorter than 1e-4 is skipped -- the overwhelmingly common case.  So 32
// consecutive steps are evaluated by 32 lanes, each rebuilding its own state from
// the points, and a warp scan orders their output; a chunk that does contain a
// skipped (degenerate) segment is replayed serially by lane 0 with the exact
// reference recurrence.  EMIT selects count-only or write mode.
template <bool EMIT>
__device__ uint32_t warp_half(const device_frame &f, uint32_t h, uint32_t out_base, uint32_t loop_id)
{
    const int lane = threadIdx.x & 31;
    stroke_src src = f.sources[h >> 1];
    loop_span l = f.loops[src.loop];
    if (l.count < 2) return 0;
    const draw_rec &d = f.draws[src.draw_closed & 0x7fffffffu];
    const bool closed = (src.draw_closed >> 31) != 0;
    const stroke_style st = style_of(d);
    const bool backwards = (h & 1) != 0;
    const uint32_t origin = backwards ? l.first + l.count - 1 : l.first;
    const uint32_t count = l.count;
    const float eps = 1.0e-4f;
    auto point_at = [&](uint32_t step) -> vec2 {       // user-space point visited at `step`
        uint32_t k = step % count;
        return apply(st.inv, ld(f.pts, backwards ? origin - k : origin + k));
    };
    walk_state ws;
    ws.pivot = point_at(0);
    ws.tin = v2(0.0f, 0.0f);
    ws.lin = 0.0f;
    uint32_t total_steps = count;                        // grows by `sf` once a closed walk finds its first join
    uint32_t first_join = 0xffffffffu;
    uint32_t emitted = 0;
    uint32_t base = 1;                                   // step 0 revisits the start point: a no-op
    while (base < total_steps) {
        const uint32_t limit = total_steps;              // steps known when this chunk starts
        const uint32_t chunk = min(32u, limit - base);
        uint32_t s = base + uint32_t(lane);
        bool valid = s < limit;
        vec2 q = valid ? point_at(s) : v2(0.0f, 0.0f);
        vec2 prev = v2(__shfl_up_sync(0xffffffffu, q.x, 1), __shfl_up_sync(0xffffffffu, q.y, 1));
        vec2 prev2 = v2(__shfl_up_sync(0xffffffffu, q.x, 2), __shfl_up_sync(0xffffffffu, q.y, 2));
        if (lane == 0) prev = ws.pivot;
        if (lane == 1) prev2 = ws.pivot;
        bool accepted = !valid || vlen(q - prev) >= eps;
        bool fast = __all_sync(0xffffffffu, accepted);
        if (fast) {
            walk_state mine;
            mine.pivot = prev;
            if (lane == 0) { mine.tin = ws.tin; mine.lin = ws.lin; }
            else { mine.tin = unit(prev - prev2); mine.lin = vlen(prev - prev2); }
            uint32_t n = 0;
            bool joined = false;
            if (valid) {
                count_sink cs = { 0 };
                walk_state probe = mine;
                joined = join_step(probe, q, st, cs);
                n = uint32_t(cs.n);
            }
            uint32_t incl = warp_inclusive_scan(n);
            if (EMIT && valid && n) {
                point_sink ps = { f.pts, f.pt_loop, out_base + emitted + incl - n, loop_id };
                walk_state again = mine;
                join_step(again, q, st, ps);
            }
            emitted += __shfl_sync(0xffffffffu, incl, 31);
            uint32_t jm = __ballot_sync(0xffffffffu, joined);
            if (closed && first_join == 0xffffffffu && jm) {
                first_join = base + uint32_t(__ffs(int(jm)) - 1);
                total_steps = count + first_join;
            }
            // carry the state of the last valid step to the next chunk
            uint32_t vm = __ballot_sync(0xffffffffu, valid);
            int last = 31 - __clz(int(vm));
            vec2 lp = v2(__shfl_sync(0xffffffffu, q.x, last), __shfl_sync(0xffffffffu, q.y, last));
            vec2 lprev = v2(__shfl_sync(0xffffffffu, prev.x, last), __shfl_sync(0xffffffffu, prev.y, last));
            ws.pivot = lp;
            ws.tin = unit(lp - lprev);
            ws.lin = vlen(lp - lprev);
            base += chunk;
        } else {
            // exact serial replay of this chunk by lane 0
            uint32_t n_chunk = 0, fj = first_join, ts = total_steps;
            if (lane == 0) {
                for (uint32_t k = 0; k < chunk; ++k) {
                    vec2 nxt = point_at(base + k);
                    bool joined;
                    if (EMIT) {
                        point_sink ps = { f.pts, f.pt_loop, out_base + emitted + n_chunk, loop_id };
                        joined = join_step(ws, nxt, st, ps);
                        n_chunk = ps.at - (out_base + emitted);
                    } else {
                        count_sink cs = { 0 };
                        joined = join_step(ws, nxt, st, cs);
                        n_chunk += uint32_t(cs.n);
                    }
                    if (joined && closed && fj == 0xffffffffu) { fj = base + k; ts = count + fj; }
                }
            }
            emitted += __shfl_sync(0xffffffffu, n_chunk, 0);
            first_join = __shfl_sync(0xffffffffu, fj, 0);
            total_steps = __shfl_sync(0xffffffffu, ts, 0);
            ws.pivot = v2(__shfl_sync(0xffffffffu, ws.pivot.x, 0), __shfl_sync(0xffffffffu, ws.pivot.y, 0));
            ws.tin = v2(__shfl_sync(0xffffffffu, ws.tin.x, 0), __shfl_sync(0xffffffffu, ws.tin.y, 0));
            ws.lin = __shfl_sync(0xffffffffu, ws.lin, 0);
            base += chunk;
        }
    }
    if (!closed && ws.lin != 0.0f) {
        uint32_t n_cap = 0;
        if (lane == 0) {
            if (EMIT) {
                point_sink ps = { f.pts, f.pt_loop, out_base + emitted, loop_id };
                cap_end(ws, st, ps);
                n_cap = ps.at - (out_base + emitted);
            } else {
                count_sink cs = { 0 };
                cap_end(ws, st, cs);
                n_cap = uint32_t(cs.n);
            }
        }
        emitted += __shfl_sync(0xffffffffu, n_cap, 0);
    }
    return emitted;
}

// ---- unit-parallel stroking ---------------------------------------------------
// The reference walks each polyline serially, but its state before visiting point
// k is just (last accepted point, the accepted point before that): a greedy filter
// that skips points closer than 1e-4 to the last accepted one.  Almost everywhere
// that filter accepts every point, so:
//   k_stroke_visits   (thread per visited point) marks points that are too close to
//                     their predecessor and writes the default link prev[k] = k - 1;
//   k_stroke_resolve  (warp per half) skims the marks 128 at a time and replays the
//                     greedy filter serially only across marked stretches, fixing
//                     the links there;
//   every join prev[prev[k]] -> prev[k] -> k and every cap is then an independent
//   unit: count -> scan -> emit, one thread each.
// A closed polyline is visited count + 2 times (the walk runs on to its first join).
// The rare walks where that is not true (fewer than two accepted points, or the
// first join not at the second point) go to the exact serial warp_half instead.

constexpr uint32_t kNoLink = 0xffffffffu;

struct half_view {
    uint32_t origin, count, visits, units;
    bool backwards, closed;
};

__device__ __forceinline__ half_view view_half(const device_frame &f, uint32_t h, stroke_style &st)
{
    half_view v;
    stroke_src src = f.sources[h >> 1];
    loop_span l = f.loops[src.loop];
    st = style_of(f.draws[src.draw_closed & 0x7fffffffu]);
    v.closed = (src.draw_closed >> 31) != 0;
    v.backwards = (h & 1) != 0;
    v.count = l.count;
    v.origin = v.backwards ? l.first + l.count - 1 : l.first;
    v.visits = l.count < 2 ? 0u : (v.closed ? l.count + 2 : l.count);
    // units: unit i < visits - 1 is the join made on visit i + 1, the last unit is the cap slot;
    // every half owns at least one unit so that its output offset is always defined
    v.units = v.visits ? v.visits : 1u;
    return v;
}

// k-th visited point in user space
__device__ __forceinline__ vec2 visit(const device_frame &f, const half_view &v, const stroke_style &st, uint32_t k)
{
    uint32_t i = k < v.count ? k : k - v.count;
    return apply(st.inv, ld(f.pts, v.backwards ? v.origin - i : v.origin + i));
}

__device__ __forceinline__ uint32_t n_halves(const device_frame &f)
{
    return 2 * (f.n_dash_items ? f.hdr->n_sources : f.n_static_sources);
}

// units per half + in-kernel partial scan; total to hdr->n_stroke_units
__global__ void __launch_bounds__(kBlock) k_stroke_plan(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t n = hd->overflow ? 0 : n_halves(f), begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) {
        stroke_style st;
        half_view v = view_half(f, first + k, st);
        f.half_count[first + k] = v.units;
        f.half_dirty[first + k] = 0;
        sum += v.units;
    }
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[3 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + 3 * kGrid, &hd->tickets[3], &hd->n_stroke_units, sm);
}

// half_count (units per half) -> half_unit_off (exclusive), same slices as k_stroke_plan
__global__ void __launch_bounds__(kBlock) k_stroke_plan_apply(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t n = hd->overflow ? 0 : n_halves(f), begin, end, ipt;
    block_slice(n, begin, end, ipt);
    if (blockIdx.x == 0 && threadIdx.x == 0 && hd->n_stroke_units > f.cap_stroke_units) atomicOr(&hd->overflow, OVF_POINTS);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) sum += f.half_count[first + k];
    uint32_t total;
    uint32_t at = block_exclusive_scan(sum, sm, total) + f.partials[3 * kGrid + blockIdx.x];
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) {
        f.half_unit_off[first + k] = at;
        at += f.half_count[first + k];
        if (first + k == n - 1) f.half_unit_off[n] = at;
    }
}

__device__ __forceinline__ uint32_t find_half(const uint32_t *off, uint32_t n, uint32_t unit)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= unit) lo = mid; else hi = mid;
    }
    return lo;
}

// thread per visit: too-close mark and default link
__global__ void __launch_bounds__(kBlock) k_stroke_visits(device_frame f)
{
    grid_dependency_wait();
    frame_header *hd = f.hdr;
    uint32_t n = hd->overflow ? 0 : hd->n_stroke_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t nh = n_halves(f);
    uint32_t first = begin + threadIdx.x * ipt;
    if (first >= end) return;
    uint32_t h = find_half(f.half_unit_off, nh, first);
    stroke_style st;
    half_view v = view_half(f, h, st);
    for (uint32_t i = 0; i < ipt; ++i) {
        uint32_t u = first + i;
        if (u >= end) break;
        while (u >= f.half_unit_off[h + 1]) { ++h; v = view_half(f, h, st); }
        uint32_t k = u - f.half_unit_off[h];
        bool close = false;
        if (k >= 1 && k < v.visits) close = vlen(visit(f, v, st, k) - visit(f, v, st, k - 1)) < 1.0e-4f;
        f.visit_close[u] = close ? 1 : 0;
        f.visit_prev[u] = k ? k - 1 : kNoLink;
    }
}

// warp per half: replay the greedy filter across marked stretches only
__global__ void __launch_bounds__(kBlock) k_stroke_resolve(device_frame f)
{
    grid_dependency_wait();
    frame_header *hd = f.hdr;
    if (hd->overflow) return;
    uint32_t nh = n_halves(f);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t h = warp; h < nh; h += n_warps) {
        stroke_style st;
        half_view v = view_half(f, h, st);
        if (!v.visits) { if (lane == 0) f.half_last[h] = 0; continue; }
        const uint32_t off = f.half_unit_off[h];
        uint32_t pivot = 0, accepted = 0, second = kNoLink;     // second: visit of the 2nd acceptance
        bool in_sync = true;
        for (uint32_t base = 1; base < v.visits; base += 128) {
            uint32_t marks = 0;
#pragma unroll
            for (uint32_t m = 0; m < 4; ++m) {
                uint32_t k = base + 4 * lane + m;
                if (k < v.visits && f.visit_close[off + k]) marks |= 1u << m;
            }
            uint32_t n_here = min(128u, v.visits - base);
            if (in_sync && !__any_sync(0xffffffffu, marks != 0)) {       // everything accepted
                if (accepted < 2 && accepted + n_here >= 2) second = base + (1 - accepted);
                accepted += n_here;
                pivot = base + n_here - 1;
                continue;
            }
            // serial replay of this stretch by lane 0
            if (lane == 0) {
                vec2 pv = visit(f, v, st, pivot);
                for (uint32_t k = base; k < base + n_here; ++k) {
                    vec2 q = visit(f, v, st, k);
                    if (vlen(q - pv) >= 1.0e-4f) {
                        f.visit_prev[off + k] = pivot;
                        pivot = k; pv = q;
                        if (++accepted == 2) second = k;
                    } else
                        f.visit_prev[off + k] = kNoLink;
                }
            }
            pivot = __shfl_sync(0xffffffffu, pivot, 0);
            accepted = __shfl_sync(0xffffffffu, accepted, 0);
            second = __shfl_sync(0xffffffffu, second, 0);
            in_sync = pivot == base + n_here - 1;
        }
        if (lane == 0) {
            f.half_last[h] = pivot;
            // a closed walk runs exactly two visits past its end only if its first join
            // happened on visit 2; anything else is replayed by the serial fallback
            if (v.closed && second != 2) f.half_dirty[h] = 1;
        }
    }
}

// Unit i < visits - 1: the join made when the walk visits point i + 1.  Last unit: the cap.
template <class Sink>
__device__ __forceinline__ void stroke_unit(const device_frame &f, const half_view &v, const stroke_style &st,
                                            uint32_t off, uint32_t last, uint32_t unit_index, Sink &sink)
{
    if (!v.visits) return;
    walk_state ws;
    const uint32_t k = unit_index + 1;
    if (k >= v.visits) {                                 // cap of an open half, after its last accepted point
        if (v.closed || last == 0) return;
        uint32_t p = f.visit_prev[off + last];
        vec2 a = visit(f, v, st, p), b = visit(f, v, st, last);
        ws.pivot = b; ws.tin = unit(b - a); ws.lin = vlen(b - a);
        cap_end(ws, st, sink);
        return;
    }
    uint32_t p1 = f.visit_prev[off + k];
    if (p1 == kNoLink || p1 == 0) return;                // skipped point, or no incoming segment yet
    uint32_t p2 = f.visit_prev[off + p1];
    vec2 a = visit(f, v, st, p2), b = visit(f, v, st, p1);
    ws.pivot = b; ws.tin = unit(b - a); ws.lin = vlen(b - a);
    join_step(ws, visit(f, v, st, k), st, sink);
}

// points per unit
__global__ void __launch_bounds__(kBlock) k_stroke_unit_count(device_frame f)
{
    grid_dependency_wait();
    frame_header *hd = f.hdr;
    uint32_t n = hd->overflow ? 0 : hd->n_stroke_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t nh = n_halves(f);
    uint32_t first = begin + threadIdx.x * ipt;
    if (first >= end) return;
    uint32_t h = find_half(f.half_unit_off, nh, first);
    stroke_style st;
    half_view v = view_half(f, h, st);
    for (uint32_t i = 0; i < ipt; ++i) {
        uint32_t u = first + i;
        if (u >= end) break;
        while (u >= f.half_unit_off[h + 1]) { ++h; v = view_half(f, h, st); }
        count_sink cs = { 0 };
        if (!f.half_dirty[h]) stroke_unit(f, v, st, f.half_unit_off[h], f.half_last[h], u - f.half_unit_off[h], cs);
        f.stroke_unit_pts[u] = uint32_t(cs.n);
    }
}

// exact serial redo of dirty halves: all their points are booked on their first unit
__global__ void __launch_bounds__(kBlock) k_stroke_fallback_count(device_frame f)
{
    grid_dependency_wait();
    frame_header *hd = f.hdr;
    if (hd->overflow) return;
    uint32_t nh = n_halves(f);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t h = warp; h < nh; h += n_warps) {
        if (!f.half_dirty[h]) continue;
        uint32_t c = warp_half<false>(f, h, 0, 0);
        if (lane == 0) f.stroke_unit_pts[f.half_unit_off[h]] = c;
    }
}

// block sums of stroke_unit_pts -> partials -> total stroke points
__global__ void __launch_bounds__(kBlock) k_stroke_unit_sums(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t n = hd->overflow ? 0 : hd->n_stroke_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) sum += f.stroke_unit_pts[first + k];
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[6 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + 6 * kGrid, &hd->tickets[7], &hd->n_stroke_points, sm);
}

__global__ void __launch_bounds__(kBlock) k_stroke_unit_emit(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t base = hd->n_line_points + hd->n_dash_points;
    if (base + hd->n_stroke_points > f.cap_pts) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&hd->overflow, OVF_POINTS);
        return;
    }
    uint32_t n = hd->overflow ? 0 : hd->n_stroke_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t nh = n_halves(f);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) sum += f.stroke_unit_pts[first + k];
    uint32_t total;
    uint32_t at = block_exclusive_scan(sum, sm, total) + f.partials[6 * kGrid + blockIdx.x];
    if (first >= end) return;
    uint32_t h = find_half(f.half_unit_off, nh, first);
    stroke_style st;
    half_view v = view_half(f, h, st);
    for (uint32_t i = 0; i < ipt; ++i) {
        uint32_t u = first + i;
        if (u >= end) break;
        while (u >= f.half_unit_off[h + 1]) { ++h; v = view_half(f, h, st); }
        uint32_t k = u - f.half_unit_off[h], mine = f.stroke_unit_pts[u];
        if (k == 0) f.half_offset[h] = at;               // where this half's output starts
        if (u == n - 1) f.half_offset[nh] = at + mine;
        if (!f.half_dirty[h] && mine) {
            // closed source: each half is its own loop; open: both halves form one loop
            uint32_t loop_id = f.stroke_loop_base + (v.closed ? h : (h & ~1u));
            point_sink ps = { f.pts, f.pt_loop, base + at, loop_id };
            stroke_unit(f, v, st, f.half_unit_off[h], f.half_last[h], k, ps);
        }
        at += mine;
    }
}

// dirty halves written serially; loop table for all halves
__global__ void __launch_bounds__(kBlock) k_stroke_finish(device_frame f)
{
    grid_dependency_wait();
    frame_header *hd = f.hdr;
    if (hd->overflow) return;
    uint32_t nh = n_halves(f);
    uint32_t base = hd->n_line_points + hd->n_dash_points;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t h = warp; h < nh; h += n_warps) {
        bool closed = (f.sources[h >> 1].draw_closed >> 31) != 0;
        uint32_t at = f.half_offset[h], count = f.half_offset[h + 1] - at;
        if (f.half_dirty[h]) warp_half<true>(f, h, base + at, f.stroke_loop_base + (closed ? h : (h & ~1u)));
        if (lane == 0) {
            loop_span l;
            l.first = base + at;
            l.count = closed ? count : ((h & 1) == 0 ? f.half_offset[h + 2] - at : 0);
            f.loops[f.stroke_loop_base + h] = l;
        }
    }
}

__global__ void k_join_math(const float *x, uint32_t n, float *acos_out, float *tan_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        acos_out[i] = join_acosf(x[i]);
        tan_out[i] = join_tanf(x[i]);
    }
}

}  // namespace

void launch_join_math(const float *x, uint32_t n, float *acos_out, float *tan_out, cudaStream_t s)
{
    k_join_math<<<kGrid, kBlock, 0, s>>>(x, n, acos_out, tan_out);
}

void launch_glyphs(const device_frame &f, cudaStream_t s)
{
    if (!f.n_glyph_insts) return;
    const uint32_t ctas = (f.n_glyph_insts + kBlock / 32 - 1) / (kBlock / 32);
    launch_pdl(k_glyph_instances, std::min<uint32_t>(ctas, 8 * kSMs), kBlock, 0, s, f);
}

void launch_flatten(const device_frame &f, uint32_t n_units, cudaStream_t s)
{
    if (!n_units) return;
    launch_pdl(k_flatten_count, kGrid, kBlock, 0, s, f);
    launch_pdl(k_flatten_emit, kGrid, kBlock, 0, s, f);
    launch_pdl(k_subpath_loops, kGrid, kBlock, 0, s, f);
}

void launch_dash(const device_frame &f, cudaStream_t s)
{
    if (!f.n_dash_items) return;
    launch_pdl(k_dash_count, kGrid, kBlock, 0, s, f);
    launch_pdl(k_dash_emit, kGrid, kBlock, 0, s, f);
}

void launch_stroke(const device_frame &f, cudaStream_t s)
{
    if (!f.n_static_sources && !f.n_dash_items) return;
    launch_pdl(k_stroke_plan, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_plan_apply, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_visits, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_resolve, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_unit_count, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_fallback_count, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_unit_sums, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_unit_emit, kGrid, kBlock, 0, s, f);
    launch_pdl(k_stroke_finish, kGrid, kBlock, 0, s, f);
}

}  // namespace cb200
