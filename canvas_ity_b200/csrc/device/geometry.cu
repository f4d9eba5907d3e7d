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
//   K3  one thread per half stroke (source polyline x direction), same
//       count/emit structure (add_half_stroke / stroke_lines, hpp:1949-2100).
//       Round joins and caps call the K1 device routine for their arcs.
#include "frame.cuh"

namespace cb200 {

namespace {

__device__ __forceinline__ vec2 ld(const float2 *p, uint32_t i) { float2 v = p[i]; return v2(v.x, v.y); }

// ------------------------------------------------------------------- K1 ----

template <class Sink>
__device__ __forceinline__ void flatten_unit(const device_frame &f, uint32_t u, Sink &sink)
{
    unit_rec un = f.units[u];
    subpath_rec sp = f.subpaths[un.subpath];
    if (un.index == 0) { sink.put(ld(f.in_points, sp.first_point)); return; }
    uint32_t at = sp.first_point + 3 * (un.index - 1);
    flatten_cubic(ld(f.in_points, at), ld(f.in_points, at + 1), ld(f.in_points, at + 2),
                  ld(f.in_points, at + 3), f.draws[sp.draw].angular, sink);
}

__global__ void __launch_bounds__(kBlock) k_flatten_count(device_frame f)
{
    __shared__ uint32_t sm[33];
    uint32_t n = f.hdr->n_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt; ++k) {
        uint32_t u = first + k;
        if (u >= end) break;
        count_sink cs = { 0 };
        flatten_unit(f, u, cs);
        f.unit_count[u] = uint32_t(cs.n);
        sum += uint32_t(cs.n);
    }
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[blockIdx.x] = total;
    finish_partials(f.partials, &f.hdr->tickets[0], &f.hdr->n_line_points, sm);
}

struct point_sink {
    float2 *out; uint32_t *loop_of; uint32_t at, loop;
    __device__ __forceinline__ void put(vec2 p) { out[at] = make_float2(p.x, p.y); loop_of[at] = loop; ++at; }
};

__global__ void __launch_bounds__(kBlock) k_flatten_emit(device_frame f)
{
    __shared__ uint32_t sm[33];
    uint32_t n = f.hdr->n_units, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    if (f.hdr->n_line_points > f.cap_pts) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&f.hdr->overflow, OVF_POINTS);
        return;
    }
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) sum += f.unit_count[first + k];
    uint32_t total;
    uint32_t at = block_exclusive_scan(sum, sm, total) + f.partials[blockIdx.x];
    for (uint32_t k = 0; k < ipt; ++k) {
        uint32_t u = first + k;
        if (u >= end) break;
        f.unit_offset[u] = at;
        point_sink ps = { f.pts, f.pt_loop, at, f.units[u].subpath };
        flatten_unit(f, u, ps);
        at = ps.at;
        if (u == n - 1) f.unit_offset[n] = at;
    }
}

// loops[s] = point span of subpath s in the K1 output
__global__ void k_subpath_loops(device_frame f)
{
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

// One side of a polyline, walked from index `first` to `last` (either
// direction) in user space; emits device-space outline points.
template <class Sink>
__device__ void walk_half(const float2 *pts, uint32_t first, uint32_t last, bool closed,
                          const stroke_style &st, Sink &sink)
{
    const float eps = 1.0e-4f;
    vec2 tin = v2(0.0f, 0.0f);
    float lin = 0.0f;
    vec2 pivot = apply(st.inv, ld(pts, first));
    uint32_t finish = first, i = first;
    do {
        vec2 nxt = apply(st.inv, ld(pts, i));
        vec2 tout = unit(nxt - pivot);
        float lout = vlen(nxt - pivot);
        if (lin != 0.0f && lout >= eps) {
            if (closed && finish == first) finish = i;
            vec2 a = pivot + st.half * perp(tin);
            vec2 b = pivot + st.half * perp(tout);
            float turn = dot(perp(tin), tout);
            if (fabsf(turn) < eps) turn = 0.0f;
            vec2 tip = turn == 0.0f ? v2(0.0f, 0.0f) : (st.half / turn) * (tout - tin);
            bool tight = dot(tip, tin) < -lin && dot(tip, tout) > lout;
            bool wrap = turn > 0.0f && tight;        // inner join tighter than the segments
            if (wrap) {
                vec2 t = a; a = b; b = t;
                t = tin; tin = tout; tout = t;
                sink.put(apply(st.fwd, b));
                sink.put(apply(st.fwd, pivot));
                sink.put(apply(st.fwd, a));
            }
            if ((turn > 0.0f && !tight) ||
                (turn != 0.0f && st.join == 0 && dot(tip, tip) <= st.miter2))
                sink.put(apply(st.fwd, pivot + tip));
            else if (st.join == 2) {
                float cosine = dot(tin, tout);
                float angle = acosf(fminf(fmaxf(cosine, -1.0f), 1.0f));
                float k = 4.0f / 3.0f * tanf(0.25f * angle);
                sink.put(apply(st.fwd, a));
                flatten_cubic(apply(st.fwd, a), apply(st.fwd, a + (k * st.half) * tin),
                              apply(st.fwd, b - (k * st.half) * tout), apply(st.fwd, b), -1.0f, sink);
            } else {
                sink.put(apply(st.fwd, a));
                sink.put(apply(st.fwd, b));
            }
            if (wrap) {
                sink.put(apply(st.fwd, b));
                sink.put(apply(st.fwd, pivot));
                sink.put(apply(st.fwd, a));
                vec2 t = tin; tin = tout; tout = t;
            }
        }
        if (lout >= eps) { tin = tout; lin = lout; pivot = nxt; }
        i = i == last ? first : (last > first ? i + 1 : i - 1);
    } while (i != finish);
    if (closed || lin == 0.0f) return;
    vec2 ahead = st.half * tin;
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

template <class Sink>
__device__ __forceinline__ void stroke_half(const device_frame &f, uint32_t h, Sink &sink)
{
    stroke_src src = f.sources[h >> 1];
    loop_span l = f.loops[src.loop];
    if (l.count < 2) return;
    const draw_rec &d = f.draws[src.draw_closed & 0x7fffffffu];
    bool closed = (src.draw_closed >> 31) != 0;
    stroke_style st = style_of(d);
    if (h & 1) walk_half(f.pts, l.first + l.count - 1, l.first, closed, st, sink);
    else walk_half(f.pts, l.first, l.first + l.count - 1, closed, st, sink);
}

__global__ void __launch_bounds__(kBlock) k_stroke_count(device_frame f)
{
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t n = 2 * (f.n_dash_items ? hd->n_sources : f.n_static_sources), begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    bool bad = hd->overflow != 0;
    for (uint32_t k = 0; k < ipt && !bad; ++k) {
        uint32_t h = first + k;
        if (h >= end) break;
        count_sink cs = { 0 };
        stroke_half(f, h, cs);
        f.half_count[h] = uint32_t(cs.n);
        sum += uint32_t(cs.n);
    }
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[3 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + 3 * kGrid, &hd->tickets[3], &hd->n_stroke_points, sm);
}

__global__ void __launch_bounds__(kBlock) k_stroke_emit(device_frame f)
{
    __shared__ uint32_t sm[33];
    frame_header *hd = f.hdr;
    uint32_t n_src = f.n_dash_items ? hd->n_sources : f.n_static_sources;
    uint32_t n = 2 * n_src, begin, end, ipt;
    block_slice(n, begin, end, ipt);
    uint32_t base = hd->n_line_points + hd->n_dash_points;
    if (base + hd->n_stroke_points > f.cap_pts) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&hd->overflow, OVF_POINTS);
        return;
    }
    if (hd->overflow) return;
    uint32_t first = begin + threadIdx.x * ipt, sum = 0;
    for (uint32_t k = 0; k < ipt && first + k < end; ++k) sum += f.half_count[first + k];
    uint32_t total;
    uint32_t at = block_exclusive_scan(sum, sm, total) + f.partials[3 * kGrid + blockIdx.x];
    for (uint32_t k = 0; k < ipt; ++k) {
        uint32_t h = first + k;
        if (h >= end) break;
        f.half_offset[h] = at;
        bool closed = (f.sources[h >> 1].draw_closed >> 31) != 0;
        uint32_t count = f.half_count[h];
        // closed source: each half is its own loop; open: both halves form one loop
        uint32_t loop_id = f.stroke_loop_base + (closed ? h : (h & ~1u));
        point_sink ps = { f.pts, f.pt_loop, base + at, loop_id };
        stroke_half(f, h, ps);
        loop_span l;
        if (closed) { l.first = base + at; l.count = count; }
        else if ((h & 1) == 0) { l.first = base + at; l.count = count + f.half_count[h + 1]; }
        else { l.first = base + at; l.count = 0; }
        f.loops[f.stroke_loop_base + h] = l;
        at += count;
        if (h == n - 1) f.half_offset[n] = at;
    }
}

}  // namespace

void launch_flatten(const device_frame &f, uint32_t n_units, cudaStream_t s)
{
    if (!n_units) return;
    k_flatten_count<<<kGrid, kBlock, 0, s>>>(f);
    k_flatten_emit<<<kGrid, kBlock, 0, s>>>(f);
    k_subpath_loops<<<kGrid, kBlock, 0, s>>>(f);
}

void launch_dash(const device_frame &f, cudaStream_t s)
{
    if (!f.n_dash_items) return;
    k_dash_count<<<kGrid, kBlock, 0, s>>>(f);
    k_dash_emit<<<kGrid, kBlock, 0, s>>>(f);
}

void launch_stroke(const device_frame &f, cudaStream_t s)
{
    if (!f.n_static_sources && !f.n_dash_items) return;
    k_stroke_count<<<kGrid, kBlock, 0, s>>>(f);
    k_stroke_emit<<<kGrid, kBlock, 0, s>>>(f);
}

}  // namespace cb200
