// oracle_raster.cpp -- TEST INFRASTRUCTURE (the parity oracle), not product code.
//
// A sequential CPU restatement of canvas_ity's rasterise + composite hot path
// (reference src/canvas_ity.hpp, "hpp" below), consuming the SAME lowered
// cb200_frame the GPU back end receives (include/canvas_b200.h).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle.so;
// the product (canvas_ity_b200/csrc) never includes or links anything here.
//
// Parity pinning: tests/test_oracle_pinning.py replays the reference's own 76
// test call streams and the tiger (tests/golden/*.cvs) through the front end's
// lowering into this oracle and checks the result against (a) the reference's
// committed RGBA8 output, (b) the reference's 76 expected image hashes with its
// own Hamming<=5 rule (test/test.cpp:2186-2261, 2618) and, when oracle/_ref is
// built, (c) the reference's float framebuffer.
//
// Stage map (each function cites what it restates):
//   flatten()        add_bezier + add_tessellation + path_to_lines  hpp:1331-1524
//   dash()           dash_lines                                      hpp:1858-1934
//   outline_stroke() add_half_stroke + stroke_lines                  hpp:1949-2100
//   scan_convert()   add_runs + lines_to_runs                        hpp:2109-2253
//   paint()          paint_pixel                                     hpp:2265-2377
//   shadow_pass()    render_shadow                                   hpp:2395-2539
//   main_pass()      render_main                                     hpp:2551-2605
//   clip_pass()      clip                                            hpp:3057-3099
//   read/write       get_image_data / put_image_data                 hpp:3348-3408
// Build with -ffp-contract=off (see oracle/Makefile).
#include "../include/canvas_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct P { float x, y; };
static inline P mk(float x, float y) { P p = { x, y }; return p; }
static inline P add(P a, P b) { return mk(a.x + b.x, a.y + b.y); }
static inline P sub(P a, P b) { return mk(a.x - b.x, a.y - b.y); }
static inline P mul(float s, P a) { return mk(a.x * s, a.y * s); }
static inline float dotp(P a, P b) { return a.x * b.x + a.y * b.y; }
static inline P rot90(P a) { return mk(-a.y, a.x); }
static inline P between(P a, P b, float t) { return add(a, mul(t, sub(b, a))); }
static inline float norm(P a) { return sqrtf(dotp(a, a)); }
static inline P unit_of(P a) { return mul(1.0f / std::max(1.0e-6f, norm(a)), a); }
static inline float sat(float v) { return std::min(std::max(v, 0.0f), 1.0f); }

struct M { float a, b, c, d, e, f; };
static inline M mat(const float *v) { M m = { v[0], v[1], v[2], v[3], v[4], v[5] }; return m; }
static inline P xf(const M &m, P p) { return mk(m.a * p.x + m.c * p.y + m.e, m.b * p.x + m.d * p.y + m.f); }

struct C { float r, g, b, a; };
static inline C cadd(C x, C y) { C c = { x.r + y.r, x.g + y.g, x.b + y.b, x.a + y.a }; return c; }
static inline C cmul(float s, C x) { C c = { x.r * s, x.g * s, x.b * s, x.a * s }; return c; }

struct Poly { std::vector<P> pts; std::vector<std::pair<size_t, bool> > subs; };   // (count, closed)

const float kThreshold = 1.0f / 8160.0f;

// ------------------------------------------------------------ flattening ----

// hpp:1331-1387: recursive halving of one curve piece.
static void halve(std::vector<P> &out, P a, P b, P c, P d, float angular, int depth_left)
{
    P ab = sub(b, a), bc = sub(c, b), cd = sub(d, c), ad = sub(d, a);
    float ab2 = dotp(ab, ab), bc2 = dotp(bc, bc), cd2 = dotp(cd, cd);
    float ad2 = std::max(1.0e-4f, dotp(ad, ad));
    float u = sat(dotp(ab, ad) / ad2), v = sat(dotp(cd, ad) / ad2);
    P miss_b = sub(add(a, mul(u, ad)), b);
    P miss_c = sub(sub(d, mul(v, ad)), c);
    float turn = 1.0f;
    if (angular > -1.0f) {
        if (ab2 * cd2 != 0.0f) turn = dotp(ab, cd) / sqrtf(ab2 * cd2);
        else if (ab2 * bc2 != 0.0f) turn = dotp(ab, bc) / sqrtf(ab2 * bc2);
        else if (bc2 * cd2 != 0.0f) turn = dotp(bc, cd) / sqrtf(bc2 * cd2);
    }
    const float tol2 = 0.125f * 0.125f;
    if ((dotp(miss_b, miss_b) <= tol2 && dotp(miss_c, miss_c) <= tol2 && turn >= angular) ||
        depth_left == 0) {
        if (angular > -1.0f && ab2 != 0.0f) out.push_back(b);
        if (angular > -1.0f && bc2 != 0.0f) out.push_back(c);
        if (angular == -1.0f || cd2 != 0.0f) out.push_back(d);
        return;
    }
    P ab_m = between(a, b, 0.5f), bc_m = between(b, c, 0.5f), cd_m = between(c, d, 0.5f);
    P abc = between(ab_m, bc_m, 0.5f), bcd = between(bc_m, cd_m, 0.5f);
    P mid = between(abc, bcd, 0.5f);
    halve(out, a, ab_m, abc, mid, angular, depth_left - 1);
    halve(out, mid, bcd, cd_m, d, angular, depth_left - 1);
}

// hpp:1420-1434 for one coordinate.
static void derivative_roots(float qa, float qb, float qc, std::vector<float> &cuts)
{
    if (fabsf(qa) > 1.0e-4f) {
        float disc = qb * qb - 4.0f * qa * qc;
        if (disc >= 0.0f) {
            float s = qb > 0.0f ? 1.0f : -1.0f;
            float t = -qb - s * sqrtf(disc);
            float r = t / (2.0f * qa);
            cuts.push_back(r);
            cuts.push_back(qc / (qa * r));
        }
    } else if (fabsf(qb) > 1.0e-4f)
        cuts.push_back(-qc / qb);
}

// hpp:1398-1487: split at extrema / max curvature, then halve each piece.
static void flatten_one(std::vector<P> &out, P a, P b, P c, P d, float angular)
{
    P ab = sub(b, a), bc = sub(c, b), cd = sub(d, c);
    if (dotp(ab, ab) == 0.0f && dotp(cd, cd) == 0.0f) { out.push_back(d); return; }
    std::vector<float> cuts;
    cuts.push_back(0.0f);
    cuts.push_back(1.0f);
    P qa = add(mul(-9.0f, bc), mul(3.0f, sub(d, a)));
    P qb = sub(mul(6.0f, add(a, c)), mul(12.0f, b));
    P qc = mul(3.0f, ab);
    derivative_roots(qa.x, qb.x, qc.x, cuts);
    derivative_roots(qa.y, qb.y, qc.y, cuts);
    float w1 = dotp(rot90(ab), bc), w2 = dotp(rot90(ab), cd), w3 = dotp(rot90(bc), cd);
    float ka = w1 - w2 + w3, kb = -2.0f * w1 + w2;
    if (fabsf(ka) > 1.0e-4f && fabsf(kb) > 1.0e-4f) cuts.push_back(-0.5f * kb / ka);
    for (size_t i = 1; i < cuts.size(); ++i) {       // insertion sort, same NaN behaviour
        float v = cuts[i];
        size_t j = i;
        while (j > 0 && v < cuts[j - 1]) { cuts[j] = cuts[j - 1]; --j; }
        cuts[j] = v;
    }
    P start = a;
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
        float t0 = cuts[i], t1 = cuts[i + 1];
        if (!(0.0f <= t0 && t1 <= 1.0f && t0 != t1)) continue;
        float rel = t0 / t1;
        P l1 = between(a, b, t1), l2 = between(b, c, t1), l3 = between(c, d, t1);
        P m1 = between(l1, l2, t1), m2 = between(l2, l3, t1);
        P n1 = between(l1, m1, rel);
        P stop = between(m1, m2, t1);
        P h2 = between(m1, stop, rel);
        P h1 = between(n1, h2, rel);
        halve(out, start, h1, h2, stop, angular, 20);
        start = stop;
    }
}

static float angular_limit(bool stroking, float line_width)       // hpp:1498-1500
{
    float ratio = 0.125f / std::max(0.5f * line_width, 0.125f);
    return stroking ? (ratio - 2.0f) * ratio * 2.0f + 1.0f : -1.0f;
}

// hpp:1495-1524 over the lowered subpaths of one draw.
static void flatten(const cb200_frame *f, const cb200_draw &d, Poly &out)
{
    float angular = angular_limit(d.kind == CB200_STROKE, d.line_width);
    out.pts.clear();
    out.subs.clear();
    for (uint32_t s = 0; s < d.n_subpaths; ++s) {
        const cb200_subpath &sp = f->subpaths[d.first_subpath + s];
        const float *p = f->points + 2 * size_t(sp.first_point);
        size_t before = out.pts.size();
        P from = mk(p[0], p[1]);
        out.pts.push_back(from);
        for (uint32_t k = 0; k < sp.n_cubics; ++k) {
            const float *q = p + 2 + 6 * size_t(k);
            P c1 = mk(q[0], q[1]), c2 = mk(q[2], q[3]), to = mk(q[4], q[5]);
            flatten_one(out.pts, from, c1, c2, to, angular);
            from = to;
        }
        out.subs.push_back(std::make_pair(out.pts.size() - before, sp.closed != 0));
    }
}

// --------------------------------------------------------------- dashing ----

// hpp:1858-1934.  Lengths are measured in user space, cut points interpolated in
// device space; a closed subpath that starts and ends inside a dash gets its last
// dash rotated in front of its first.
static void dash(const Poly &in, const float *pattern, size_t n_pattern, float dash_offset,
                 const M &inverse, Poly &out)
{
    out.pts.clear();
    out.subs.clear();
    float total = 0.0f;
    for (size_t i = 0; i < n_pattern; ++i) total += pattern[i];
    float phase = fmodf(dash_offset, total);
    if (phase < 0.0f) phase += total;
    size_t first_seg = 0;
    while (phase >= pattern[first_seg]) {
        phase -= pattern[first_seg];
        first_seg = first_seg + 1 < n_pattern ? first_seg + 1 : 0;
    }
    size_t base = 0;
    for (size_t s = 0; s < in.subs.size(); ++s) {
        size_t n = in.subs[s].first;
        bool closed = in.subs[s].second;
        size_t piece_start = out.pts.size();
        size_t seg = first_seg;
        bool on = (first_seg & 1) == 0;
        size_t first_piece_pt = out.pts.size(), first_piece_sub = out.subs.size();
        bool began_on = on;
        float until = pattern[first_seg] - phase;
        size_t i = base;
        for (; i + 1 < base + n; ++i) {
            P a = in.pts[i], b = in.pts[i + 1];
            if (on) out.pts.push_back(a);
            float len = norm(sub(xf(inverse, b), xf(inverse, a)));
            while (until < len) {
                out.pts.push_back(between(a, b, until / len));
                if (on) {
                    out.subs.push_back(std::make_pair(out.pts.size() - piece_start, false));
                    piece_start = out.pts.size();
                }
                seg = seg + 1 < n_pattern ? seg + 1 : 0;
                on = !on;
                until += pattern[seg];
            }
            until -= len;
        }
        if (on) {
            out.pts.push_back(in.pts[i]);
            out.subs.push_back(std::make_pair(out.pts.size() - piece_start, false));
            if (closed && began_on) {
                if (out.subs.size() == first_piece_sub + 1)
                    out.subs.back().second = true;           // one dash covers the loop
                else {
                    size_t tail = out.subs.back().first;
                    std::rotate(out.pts.begin() + ptrdiff_t(first_piece_pt),
                                out.pts.end() - ptrdiff_t(tail), out.pts.end());
                    out.subs[first_piece_sub].first += tail;
                    out.subs.pop_back();
                }
            }
        }
        base += n;
    }
}

// ------------------------------------------------------ stroke expansion ----

struct StrokeStyle { float half, miter2; uint32_t cap, join; M fwd, inv; };

// hpp:1949-2058: one side of one polyline, walked from `first` to `last`
// (either direction), in user space.
static void half_outline(const Poly &src, size_t first, size_t last, bool closed,
                         const StrokeStyle &st, std::vector<P> &out)
{
    P dir_in = mk(0.0f, 0.0f);
    float len_in = 0.0f;
    P at = xf(st.inv, src.pts[first]);
    size_t stop_at = first, i = first;
    do {
        P nxt = xf(st.inv, src.pts[i]);
        P dir_out = unit_of(sub(nxt, at));
        float len_out = norm(sub(nxt, at));
        if (len_in != 0.0f && len_out >= 1.0e-4f) {
            if (closed && stop_at == first) stop_at = i;
            P side_a = add(at, mul(st.half, rot90(dir_in)));
            P side_b = add(at, mul(st.half, rot90(dir_out)));
            float bend = dotp(rot90(dir_in), dir_out);
            if (fabsf(bend) < 1.0e-4f) bend = 0.0f;
            P tip = bend == 0.0f ? mk(0.0f, 0.0f) : mul(st.half / bend, sub(dir_out, dir_in));
            bool tight = dotp(tip, dir_in) < -len_in && dotp(tip, dir_out) > len_out;
            bool wrap = bend > 0.0f && tight;
            if (wrap) {
                std::swap(side_a, side_b);
                std::swap(dir_in, dir_out);
                out.push_back(xf(st.fwd, side_b));
                out.push_back(xf(st.fwd, at));
                out.push_back(xf(st.fwd, side_a));
            }
            if ((bend > 0.0f && !tight) ||
                (bend != 0.0f && st.join == 0 && dotp(tip, tip) <= st.miter2))
                out.push_back(xf(st.fwd, add(at, tip)));
            else if (st.join == 2) {
                float cs = dotp(dir_in, dir_out);
                float ang = acosf(std::min(std::max(cs, -1.0f), 1.0f));
                float k = 4.0f / 3.0f * tanf(0.25f * ang);
                out.push_back(xf(st.fwd, side_a));
                flatten_one(out, xf(st.fwd, side_a),
                            xf(st.fwd, add(side_a, mul(k * st.half, dir_in))),
                            xf(st.fwd, sub(side_b, mul(k * st.half, dir_out))),
                            xf(st.fwd, side_b), -1.0f);
            } else {
                out.push_back(xf(st.fwd, side_a));
                out.push_back(xf(st.fwd, side_b));
            }
            if (wrap) {
                out.push_back(xf(st.fwd, side_b));
                out.push_back(xf(st.fwd, at));
                out.push_back(xf(st.fwd, side_a));
                std::swap(dir_in, dir_out);
            }
        }
        if (len_out >= 1.0e-4f) {
            dir_in = dir_out;
            len_in = len_out;
            at = nxt;
        }
        i = i == last ? first : last > first ? i + 1 : i - 1;
    } while (i != stop_at);
    if (closed || len_in == 0.0f) return;
    P fwd_half = mul(st.half, dir_in);
    P side = rot90(fwd_half);
    if (st.cap == 0) {
        out.push_back(xf(st.fwd, add(at, side)));
        out.push_back(xf(st.fwd, sub(at, side)));
    } else if (st.cap == 1) {
        out.push_back(xf(st.fwd, add(add(at, fwd_half), side)));
        out.push_back(xf(st.fwd, sub(add(at, fwd_half), side)));
    } else if (st.cap == 2) {
        const float k = 0.55228475f;
        out.push_back(xf(st.fwd, add(at, side)));
        flatten_one(out, xf(st.fwd, add(at, side)),
                    xf(st.fwd, add(add(at, side), mul(k, fwd_half))),
                    xf(st.fwd, add(add(at, fwd_half), mul(k, side))),
                    xf(st.fwd, add(at, fwd_half)), -1.0f);
        flatten_one(out, xf(st.fwd, add(at, fwd_half)),
                    xf(st.fwd, sub(add(at, fwd_half), mul(k, side))),
                    xf(st.fwd, add(sub(at, side), mul(k, fwd_half))),
                    xf(st.fwd, sub(at, side)), -1.0f);
    }
}

// hpp:2070-2100 (dashing is applied by the caller).
static void outline_stroke(const Poly &src, const StrokeStyle &st, Poly &out)
{
    out.pts.clear();
    out.subs.clear();
    size_t base = 0;
    for (size_t s = 0; s < src.subs.size(); ++s) {
        size_t n = src.subs[s].first;
        bool closed = src.subs[s].second;
        size_t lo = base, hi = base + n;
        base = hi;
        if (n < 2) continue;
        size_t mark = out.pts.size();
        half_outline(src, lo, hi - 1, closed, st, out.pts);
        if (closed) {
            out.subs.push_back(std::make_pair(out.pts.size() - mark, true));
            mark = out.pts.size();
        }
        half_outline(src, hi - 1, lo, closed, st, out.pts);
        out.subs.push_back(std::make_pair(out.pts.size() - mark, true));
    }
}

// -------------------------------------------------------- scan conversion ----

struct Delta { uint16_t x; float d; };
struct Coverage {                       // sparse signed-area runs bucketed by row
    int rows = 0;
    std::vector<std::vector<Delta> > row;
    int min_x = 1 << 30, max_x = -1, min_y = 1 << 30, max_y = -1;
    bool empty() const { return max_x < 0; }
};

// hpp:2109-2170: signed trapezoid areas of one edge, pixel by pixel.
static void edge_deltas(Coverage &cov, P a, P b)
{
    if (fabsf(b.y - a.y) < 2.0e-5f) return;
    float wind = b.y > a.y ? 1.0f : -1.0f;
    if (a.x > b.x) std::swap(a, b);                       // walk left to right
    P cur = a;
    float px = floorf(cur.x), py = floorf(cur.y);
    float gx = px + 1.0f, gy = py + (b.y > a.y ? 1.0f : 0.0f);
    float dxdy = (b.x - a.x) / (b.y - a.y), dydx = (b.y - a.y) / (b.x - a.x);
    P hit_x = (b.x - a.x < 2.0e-5f) ? b : mk(gx, cur.y + (gx - cur.x) * dydx);
    P hit_y = mk(cur.x + (gy - cur.y) * dxdy, gy);
    if ((a.y < b.y && b.y < hit_y.y) || (a.y > b.y && b.y > hit_y.y)) hit_y = b;
    float ystep = b.y > a.y ? 1.0f : -1.0f;
    auto emit = [&](float fx, float fy, float v) {
        uint16_t ux = static_cast<uint16_t>(fx), uy = static_cast<uint16_t>(fy);
        if (int(uy) >= cov.rows) return;                  // never produced after clamping
        Delta e = { ux, v };
        cov.row[uy].push_back(e);
    };
    do {
        float carry = 0.0f;
        while (hit_x.x < hit_y.x) {
            float h = sat((hit_x.y - cur.y) * ystep);
            float mid = (hit_x.x + cur.x) * 0.5f;
            float area = (mid - px) * h;
            emit(px, py, (carry + h - area) * wind);
            carry = area;
            cur = hit_x;
            hit_x.x += 1.0f;
            hit_x.y = (hit_x.x - a.x) * dydx + a.y;
            px += 1.0f;
        }
        float h = sat((hit_y.y - cur.y) * ystep);
        float mid = (hit_y.x + cur.x) * 0.5f;
        float area = (mid - px) * h;
        emit(px, py, (carry + h - area) * wind);
        emit(px + 1.0f, py, area * wind);
        cur = hit_y;
        hit_y.y += ystep;
        hit_y.x = (hit_y.y - a.y) * dxdy + a.x;
        py += ystep;
        if ((a.y < b.y && b.y < hit_y.y) || (a.y > b.y && b.y > hit_y.y)) hit_y = b;
    } while (cur.y != b.y);
}

// hpp:2193-2240: Sutherland-Hodgman clip of every closed loop against the padded
// canvas, clamp, per-edge deltas (rows hold the raw runs in emission order).
static void scan_convert_raw(const Poly &poly, P offset, int width_px, int height_px, Coverage &cov)
{
    float w = float(width_px), h = float(height_px);
    cov = Coverage();
    cov.rows = height_px + 2;
    cov.row.assign(size_t(cov.rows), std::vector<Delta>());
    std::vector<P> ring, next;
    size_t base = 0;
    for (size_t s = 0; s < poly.subs.size(); ++s) {
        size_t n = poly.subs[s].first;
        ring.clear();
        for (size_t i = 0; i < n; ++i) ring.push_back(add(offset, poly.pts[base + i]));
        base += n;
        for (int side = 0; side < 4; ++side) {
            P nrm = mk(side == 0 ? 1.0f : side == 2 ? -1.0f : 0.0f,
                       side == 1 ? 1.0f : side == 3 ? -1.0f : 0.0f);
            float where = side == 2 ? w : side == 3 ? h : 0.0f;
            next.clear();
            for (size_t i = 0; i < ring.size(); ++i) {
                P a = ring[(i ? i : ring.size()) - 1], b = ring[i];
                float da = dotp(a, nrm) + where, db = dotp(b, nrm) + where;
                if (da * db < 0.0f) next.push_back(between(a, b, da / (da - db)));
                if (db >= 0.0f) next.push_back(b);
            }
            ring.swap(next);
        }
        for (size_t i = 0; i < ring.size(); ++i) {
            P a = ring[(i ? i : ring.size()) - 1], b = ring[i];
            edge_deltas(cov, mk(std::min(std::max(a.x, 0.0f), w), std::min(std::max(a.y, 0.0f), h)),
                        mk(std::min(std::max(b.x, 0.0f), w), std::min(std::max(b.y, 0.0f), h)));
        }
    }
}

// hpp:2193-2253: scan_convert_raw, then the sort and the merge of equal pixels.
static void scan_convert(const Poly &poly, P offset, int width_px, int height_px, Coverage &cov)
{
    scan_convert_raw(poly, offset, width_px, height_px, cov);
    // Merge in global (y, x, |delta|) order exactly like hpp:2244-2252: a run opens
    // a new entry only if its own delta is non-zero (the very first run always
    // does); later runs on the same pixel add to it.  The bounding box that
    // render_shadow derives (hpp:2413-2419) is over these kept entries only.
    cov.min_x = cov.min_y = 1 << 30;
    cov.max_x = cov.max_y = -1;
    bool any_kept = false;
    for (size_t y = 0; y < cov.row.size(); ++y) {
        std::vector<Delta> &r = cov.row[y];
        if (r.empty()) continue;
        std::stable_sort(r.begin(), r.end(), [](const Delta &p, const Delta &q) {
            return p.x != q.x ? p.x < q.x : fabsf(p.d) < fabsf(q.d);
        });
        size_t kept = 0;                                  // entries of this row kept so far
        for (size_t i = 0; i < r.size(); ++i) {
            if (kept && r[kept - 1].x == r[i].x) r[kept - 1].d += r[i].d;
            else if (!any_kept || r[i].d != 0.0f) { r[kept++] = r[i]; any_kept = true; }
        }
        r.resize(kept);
        for (size_t i = 0; i < kept; ++i) {
            cov.min_x = std::min(cov.min_x, int(r[i].x)); cov.max_x = std::max(cov.max_x, int(r[i].x));
            cov.min_y = std::min(cov.min_y, int(y)); cov.max_y = std::max(cov.max_y, int(y));
        }
    }
}

// Dense coverage of one row: min(|running sum|, 1) for x in [0, width).
static void row_coverage(const std::vector<Delta> &r, int width, std::vector<float> &out)
{
    out.assign(size_t(width), 0.0f);
    float sum = 0.0f;
    size_t i = 0;
    for (int x = 0; x < width; ++x) {
        while (i < r.size() && int(r[i].x) <= x) sum += r[i++].d;
        out[size_t(x)] = std::min(fabsf(sum), 1.0f);
    }
}

// ----------------------------------------------------------------- brush ----

struct Brush {
    uint32_t type = 0, flags = 0, repetition = 0;
    std::vector<C> colors;
    std::vector<float> stops;
    P start = {0, 0}, end = {0, 0};
    float r0 = 0, r1 = 0;
    int w = 0, h = 0;
};

static float srgb_to_linear(float v) { return v < 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f); }
static float linear_to_srgb(float v) { return v < 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f; }

static C decode_texel(const uint8_t *t)                 // hpp:2856-2859, 3402-3406
{
    float a = t[3] / 255.0f;
    C c = { srgb_to_linear(t[0] / 255.0f) * a, srgb_to_linear(t[1] / 255.0f) * a,
            srgb_to_linear(t[2] / 255.0f) * a, a };
    return c;
}

static Brush load_brush(const cb200_frame *f, uint32_t index)
{
    Brush b;
    const cb200_brush &s = f->brushes[index];
    b.type = s.type; b.flags = s.flags; b.repetition = s.repetition;
    b.start = mk(s.start[0], s.start[1]); b.end = mk(s.end[0], s.end[1]);
    b.r0 = s.start_radius; b.r1 = s.end_radius;
    if (s.type == CB200_BRUSH_PATTERN) {
        const cb200_image &im = f->images[s.image];
        b.w = im.width; b.h = im.height;
        const uint8_t *t = f->texels + im.texel_offset;
        b.colors.resize(size_t(b.w) * size_t(b.h));
        for (size_t i = 0; i < b.colors.size(); ++i) b.colors[i] = decode_texel(t + 4 * i);
    } else
        for (uint32_t i = 0; i < s.n_colors; ++i) {
            const float *c = f->colors + 4 * size_t(s.first_color + i);
            C col = { c[0], c[1], c[2], c[3] };
            b.colors.push_back(col);
            b.stops.push_back(f->stops[s.first_color + i]);
        }
    return b;
}

static float keys_weight(float t)                       // Catmull-Rom / Keys a = -0.5
{
    return t < 1.0f ? (1.5f * t - 2.5f) * t * t + 1.0f : ((-0.5f * t + 2.5f) * t - 4.0f) * t + 2.0f;
}

// hpp:2265-2377.
static C paint(const Brush &b, const M &inv, P device_point)
{
    C none = { 0, 0, 0, 0 };
    if (b.colors.empty()) return none;
    if (b.type == CB200_BRUSH_COLOR) return b.colors[0];
    P p = xf(inv, device_point);
    if (b.type == CB200_BRUSH_PATTERN) {
        float w = float(b.w), h = float(b.h);
        if (((b.repetition & 2) && (p.x < 0.0f || w <= p.x)) ||
            ((b.repetition & 1) && (p.y < 0.0f || h <= p.y)))
            return none;
        float sx = fabsf(inv.a) + fabsf(inv.c), sy = fabsf(inv.b) + fabsf(inv.d);
        sx = std::max(1.0f, std::min(sx, w * 0.25f));
        sy = std::max(1.0f, std::min(sy, h * 0.25f));
        float rx = 1.0f / sx, ry = 1.0f / sy;
        p = sub(p, mk(0.5f, 0.5f));
        int x0 = int(ceilf(p.x - sx * 2.0f)), y0 = int(ceilf(p.y - sy * 2.0f));
        int x1 = int(ceilf(p.x + sx * 2.0f)), y1 = int(ceilf(p.y + sy * 2.0f));
        C acc = none;
        float wsum = 0.0f;
        bool clamp_mode = (b.flags & CB200_BRUSH_CLAMP) != 0;
        for (int ty = y0; ty < y1; ++ty) {
            float wy = keys_weight(fabsf(ry * (float(ty) - p.y)));
            int yy = ty % b.h;
            if (yy < 0) yy += b.h;
            if (clamp_mode) yy = std::min(std::max(ty, 0), b.h - 1);
            for (int tx = x0; tx < x1; ++tx) {
                float wx = keys_weight(fabsf(rx * (float(tx) - p.x)));
                int xx = tx % b.w;
                if (xx < 0) xx += b.w;
                if (clamp_mode) xx = std::min(std::max(tx, 0), b.w - 1);
                float wgt = wx * wy;
                acc = cadd(acc, cmul(wgt, b.colors[size_t(yy) * size_t(b.w) + size_t(xx)]));
                wsum += wgt;
            }
        }
        return cmul(1.0f / wsum, acc);
    }
    P rel = sub(p, b.start), axis = sub(b.end, b.start);
    float along = dotp(rel, axis), axis2 = dotp(axis, axis);
    float t;
    if (b.type == CB200_BRUSH_LINEAR) {
        if (axis2 == 0.0f) return none;
        t = along / axis2;
    } else {
        float dr = b.r1 - b.r0;
        float qa = axis2 - dr * dr;
        float qb = -2.0f * (along + b.r0 * dr);
        float qc = dotp(rel, rel) - b.r0 * b.r0;
        float disc = qb * qb - 4.0f * qa * qc;
        if (disc < 0.0f || (axis2 == 0.0f && dr == 0.0f)) return none;
        float root = sqrtf(disc), inv2a = 1.0f / (2.0f * qa);
        float ta = (-qb - root) * inv2a, tb = (-qb + root) * inv2a;
        if (b.r0 + dr * tb >= 0.0f) t = tb;
        else if (b.r0 + dr * ta >= 0.0f) t = ta;
        else return none;
    }
    size_t hi = size_t(std::upper_bound(b.stops.begin(), b.stops.end(), t) - b.stops.begin());
    C c;
    if (hi == 0) c = b.colors.front();
    else if (hi == b.stops.size()) c = b.colors.back();
    else {
        float m = (t - b.stops[hi - 1]) / (b.stops[hi] - b.stops[hi - 1]);
        C lo = b.colors[hi - 1], up = b.colors[hi];
        C step = { up.r - lo.r, up.g - lo.g, up.b - lo.b, up.a - lo.a };
        c = cadd(lo, cmul(m, step));
    }
    C pm = { c.r * c.a, c.g * c.a, c.b * c.a, c.a };
    return pm;
}

// --------------------------------------------------------------- canvas ----

struct Canvas {
    int w, h;
    std::vector<C> fb;
    std::map<uint32_t, std::vector<float> > masks;      // dense visibility planes
    float vis(uint32_t slot, int x, int y) const
    {
        if (slot == 0) return 1.0f;
        std::map<uint32_t, std::vector<float> >::const_iterator it = masks.find(slot);
        return it == masks.end() ? 0.0f : it->second[size_t(y) * size_t(w) + size_t(x)];
    }
};

// The 4-bit mix program of hpp:2583-2591.
static inline void blend(C &back, C fore, int op, float vis)
{
    float mf = (op & 1) ? back.a : 0.0f;
    if (op & 2) mf = 1.0f - mf;
    float mb = (op & 4) ? fore.a : 0.0f;
    if (op & 8) mb = 1.0f - mb;
    C mixed = cadd(cmul(mf, fore), cmul(mb, back));
    mixed.a = std::min(mixed.a, 1.0f);
    back = cadd(cmul(vis, mixed), cmul(1.0f - vis, back));
}

// hpp:2551-2605 with the run/mask merge walk replaced by its per-pixel meaning:
// every pixel with visibility >= 1/8160 and (coverage >= 1/8160 or a clearing op).
static void main_pass(Canvas &cv, const Coverage &cov, const Brush &br, const cb200_draw &d)
{
    int op = int(d.op);
    bool everywhere = (~op & 8) != 0;
    M inv = mat(d.inverse);
    std::vector<float> line;
    static const std::vector<Delta> no_runs;
    for (int y = 0; y < cv.h; ++y) {
        const std::vector<Delta> &r = size_t(y) < cov.row.size() ? cov.row[size_t(y)] : no_runs;
        if (r.empty() && !everywhere) continue;
        row_coverage(r, cv.w, line);
        for (int x = 0; x < cv.w; ++x) {
            float c = line[size_t(x)];
            float v = std::min(fabsf(cv.vis(d.mask_src, x, y)), 1.0f);
            if (!((c >= kThreshold || everywhere) && v >= kThreshold)) continue;
            C fore = cmul(c * d.global_alpha, paint(br, inv, mk(float(x) + 0.5f, float(y) + 0.5f)));
            blend(cv.fb[size_t(y) * size_t(cv.w) + size_t(x)], fore, op, v);
        }
    }
}

// One extended-box pass over `n` samples with stride `step` (hpp:2460-2503): the
// reference's running-sum formulation, zero outside [0, n).
static void box_pass(float *data, size_t n, size_t step, size_t radius, float w1, float w2,
                     std::vector<float> &tmp)
{
    tmp.resize(n + radius + 2);
    for (size_t i = 0; i < n; ++i) tmp[i] = data[i * step];
    // the reference reads tmp[radius+1] and tmp[0..radius] unguarded: its scratch is
    // max(width,height) long and already holds the previous line there; inside the
    // working area those indexes are < n whenever n > radius+1, which the border
    // (3*(radius+1) on each side) guarantees.
    float run = w1 * tmp[radius + 1];
    for (size_t i = 0; i <= radius; ++i) run += (w1 + w2) * tmp[i];
    data[0] = run;
    for (size_t i = 1; i < n; ++i) {
        if (i >= radius + 1) run -= w2 * tmp[i - radius - 1];
        if (i >= radius + 2) run -= w1 * tmp[i - radius - 2];
        if (i + radius < n) run += w2 * tmp[i + radius];
        if (i + radius + 1 < n) run += w1 * tmp[i + radius + 1];
        data[i * step] = run;
    }
}

// hpp:2395-2539.
static void shadow_pass(Canvas &cv, const Poly &poly, const Brush &br, const cb200_draw &d)
{
    if (d.shadow_color[3] == 0.0f ||
        (d.shadow_blur == 0.0f && d.shadow_offset_x == 0.0f && d.shadow_offset_y == 0.0f))
        return;
    float sigma2 = 0.25f * d.shadow_blur * d.shadow_blur;
    size_t radius = size_t(0.5f * sqrtf(4.0f * sigma2 + 1.0f) - 0.5f);
    int border = 3 * (int(radius) + 1);
    P off = mk(float(border) + d.shadow_offset_x, float(border) + d.shadow_offset_y);
    Coverage cov;
    scan_convert(poly, off, cv.w + 2 * border, cv.h + 2 * border, cov);
    int left = cv.w + 2 * border, right = 0, top = cv.h + 2 * border, bottom = 0;
    if (!cov.empty()) {
        left = std::min(left, cov.min_x); right = std::max(right, cov.max_x);
        top = std::min(top, cov.min_y); bottom = std::max(bottom, cov.max_y);
    }
    left = std::max(left - border, 0);
    right = std::min(right + border, cv.w + 2 * border) + 1;
    top = std::max(top - border, 0);
    bottom = std::min(bottom + border, cv.h + 2 * border);
    size_t bw = size_t(std::max(right - left, 0)), bh = size_t(std::max(bottom - top, 0));
    std::vector<float> plane(bw * bh + std::max(bw, bh) + radius + 4, 0.0f);
    M inv = mat(d.inverse);
    // hpp:2430-2452: between two runs of a row the pixels get the running coverage; after the LAST run of a row only
    // that run's own pixel does (to = x + 1 when the next run starts another row) -- unlike render_main, no mask run
    // at the right canvas edge carries a residual sum onward -- and the last run of all is never painted (the loop
    // paints on arrival of the next run).
    int last_row = -1;
    for (int y = top; y < bottom; ++y)
        if (size_t(y) < cov.row.size() && !cov.row[size_t(y)].empty()) last_row = y;
    for (int y = top; y < bottom; ++y) {
        if (size_t(y) >= cov.row.size() || cov.row[size_t(y)].empty()) continue;
        const std::vector<Delta> &r = cov.row[size_t(y)];
        float sum = 0.0f;
        for (size_t i = 0; i < r.size(); ++i) {
            sum += r[i].d;
            const bool last_of_row = i + 1 == r.size();
            if (last_of_row && y == last_row) break;
            int x0 = int(r[i].x), x1 = last_of_row ? x0 + 1 : int(r[i + 1].x);
            float c = std::min(fabsf(sum), 1.0f);
            if (c < kThreshold) continue;
            for (int x = std::max(x0, left); x < std::min(x1, right); ++x) {
                P centre = sub(mk(float(x) + 0.5f, float(y) + 0.5f), off);
                plane[size_t(y - top) * bw + size_t(x - left)] = c * paint(br, inv, centre).a;
            }
        }
    }
    float alpha = float(2 * radius + 1) * (float(radius * (radius + 1)) - sigma2) /
                  (2.0f * sigma2 - float(6 * (radius + 1) * (radius + 1)));
    float div = 2.0f * (alpha + float(radius)) + 1.0f;
    float w1 = alpha / div, w2 = (1.0f - alpha) / div;
    std::vector<float> tmp;
    if (bw > 0 && bh > 0) {
        for (size_t y = 0; y < bh; ++y)
            for (int pass = 0; pass < 3; ++pass) box_pass(&plane[y * bw], bw, 1, radius, w1, w2, tmp);
        for (size_t x = 0; x < bw; ++x)
            for (int pass = 0; pass < 3; ++pass) box_pass(&plane[x], bh, bw, radius, w1, w2, tmp);
    }
    int op = int(d.op);
    C tint = { d.shadow_color[0], d.shadow_color[1], d.shadow_color[2], d.shadow_color[3] };
    for (int y = 0; y < cv.h; ++y) {
        if (!(top <= y + border && y + border < bottom)) continue;
        for (int x = std::max(0, left - border); x < std::min(cv.w, right - border); ++x) {
            float v = std::min(fabsf(cv.vis(d.mask_src, x, y)), 1.0f);
            if (v < kThreshold) continue;
            float s = plane[size_t(y + border - top) * bw + size_t(x + border - left)];
            blend(cv.fb[size_t(y) * size_t(cv.w) + size_t(x)], cmul(d.global_alpha * s, tint), op, v);
        }
    }
}

// hpp:3057-3099 as a dense product of clamped coverages.
static void clip_pass(Canvas &cv, const Coverage &cov, const cb200_draw &d)
{
    std::vector<float> plane(size_t(cv.w) * size_t(cv.h), 0.0f);
    std::vector<float> line;
    for (int y = 0; y < cv.h; ++y) {
        if (size_t(y) >= cov.row.size() || cov.row[size_t(y)].empty()) continue;
        row_coverage(cov.row[size_t(y)], cv.w, line);
        for (int x = 0; x < cv.w; ++x)
            plane[size_t(y) * size_t(cv.w) + size_t(x)] =
                line[size_t(x)] * std::min(fabsf(cv.vis(d.mask_src, x, y)), 1.0f);
    }
    cv.masks[d.mask_dst].swap(plane);
}

static void run_draw(Canvas &cv, const cb200_frame *f, const cb200_draw &d)
{
    Poly lines, work;
    flatten(f, d, lines);
    if (d.kind == CB200_STROKE) {
        if (d.n_dash) {
            dash(lines, f->dashes + d.first_dash, d.n_dash, d.dash_offset, mat(d.inverse), work);
            lines.pts.swap(work.pts);
            lines.subs.swap(work.subs);
        }
        StrokeStyle st;
        st.half = d.line_width * 0.5f;
        st.miter2 = d.miter_limit * d.miter_limit * st.half * st.half;
        st.cap = d.cap; st.join = d.join;
        st.fwd = mat(d.forward); st.inv = mat(d.inverse);
        outline_stroke(lines, st, work);
        lines.pts.swap(work.pts);
        lines.subs.swap(work.subs);
    }
    Coverage cov;
    if (d.kind == CB200_CLIP) {
        scan_convert(lines, mk(0.0f, 0.0f), cv.w, cv.h, cov);
        clip_pass(cv, cov, d);
        return;
    }
    Brush br = load_brush(f, d.brush);
    shadow_pass(cv, lines, br, d);
    scan_convert(lines, mk(0.0f, 0.0f), cv.w, cv.h, cov);
    main_pass(cv, cov, br, d);
}

}  // namespace

// ------------------------------------------------------------------ C API ----

extern "C" {

void *oracle_canvas_create(int width, int height)
{
    Canvas *cv = new Canvas;
    cv->w = width;
    cv->h = height;
    C zero = { 0, 0, 0, 0 };
    cv->fb.assign(size_t(width) * size_t(height), zero);
    return cv;
}

void oracle_canvas_destroy(void *canvas) { delete static_cast<Canvas *>(canvas); }

// Glyph instances (cb200_glyph_inst): what the reference's add_glyph (hpp:1533-1696) emits for
// one glyph under the matrix of text_to_lines (hpp:1793-1846), from the parsed outline: on-curve
// points as they are, the implied midpoint between two off-curve points, quadratic pieces around
// an off-curve point (kept as degree-elevated cubics for flatten()), straight pieces as (from, to, to).
static void expand_glyphs(const cb200_frame *f, std::vector<float> &pts)
{
    pts.assign(f->points, f->points + 2 * size_t(f->n_points));
    pts.resize(2 * (size_t(f->n_points) + f->n_glyph_points), 0.0f);
    float *region = pts.data() + 2 * size_t(f->n_points);
    for (uint32_t g = 0; g < f->n_glyphs; ++g) {
        const cb200_glyph_inst &gi = f->glyphs[g];
        const cb200_glyph_atlas &at = f->atlases[gi.atlas];
        const cb200_glyph_outline &o = at.outlines[gi.outline];
        const M m = mat(gi.m);
        float *dst = region + 2 * size_t(gi.first_point);
        auto point = [&](uint32_t k) { const float *u = at.points + 2 * size_t(o.first_point + k); return xf(m, mk(u[0], u[1])); };
        auto end_point = [&](uint32_t a, uint32_t b) { return a == b ? point(a) : between(point(a), point(b), 0.5f); };
        auto put = [&](uint32_t slot, P p) { dst[2 * size_t(slot)] = p.x; dst[2 * size_t(slot) + 1] = p.y; };
        for (uint32_t k = 0; k < o.n_segs; ++k) {
            const cb200_glyph_seg &sg = at.segs[o.first_seg + k];
            P from = end_point(sg.from_a, sg.from_b), to = end_point(sg.to_a, sg.to_b);
            P c1 = from, c2 = to;
            if (!(sg.flags & CB200_SEG_LINE)) {
                P c = point(sg.ctrl);
                c1 = between(from, c, 2.0f / 3.0f);
                c2 = between(to, c, 2.0f / 3.0f);
            }
            if (sg.flags & CB200_SEG_FIRST) put(sg.out - 1, from);
            put(sg.out, c1);
            put(sg.out + 1, c2);
            put(sg.out + 2, to);
        }
    }
}

void oracle_submit(void *canvas, const cb200_frame *frame)
{
    Canvas &cv = *static_cast<Canvas *>(canvas);
    cb200_frame local = *frame;
    std::vector<float> pts;
    std::vector<cb200_subpath> subs;
    if (frame->n_glyphs) {                       // text drawn as glyph instances: expand them first
        expand_glyphs(frame, pts);
        subs.assign(frame->subpaths, frame->subpaths + frame->n_subpaths);
        for (size_t i = 0; i < subs.size(); ++i)
            if (subs[i].instanced) { subs[i].first_point += frame->n_points; subs[i].instanced = 0; }
        local.points = pts.data();
        local.n_points = frame->n_points + frame->n_glyph_points;
        local.subpaths = subs.data();
        local.n_glyphs = 0;
    }
    for (uint32_t i = 0; i < local.n_draws; ++i) run_draw(cv, &local, local.draws[i]);
}

// get_image_data, hpp:3348-3381.
void oracle_read_rgba8(void *canvas, uint8_t *dst, int width, int height, int stride, int x, int y)
{
    Canvas &cv = *static_cast<Canvas *>(canvas);
    static const float bayer[4][4] = {
        { 0.5f / 16, 8.5f / 16, 2.5f / 16, 10.5f / 16 }, { 12.5f / 16, 4.5f / 16, 14.5f / 16, 6.5f / 16 },
        { 3.5f / 16, 11.5f / 16, 1.5f / 16, 9.5f / 16 }, { 15.5f / 16, 7.5f / 16, 13.5f / 16, 5.5f / 16 } };
    for (int j = 0; j < height; ++j)
        for (int i = 0; i < width; ++i) {
            int cx = x + i, cy = y + j;
            C c = { 0, 0, 0, 0 };
            if (0 <= cx && cx < cv.w && 0 <= cy && cy < cv.h) c = cv.fb[size_t(cy) * size_t(cv.w) + size_t(cx)];
            if (c.a < kThreshold) { c.r = c.g = c.b = c.a = 0.0f; }
            else { float k = 1.0f / c.a; c.r = k * c.r; c.g = k * c.g; c.b = k * c.b; }
            float t = bayer[cy & 3][cx & 3];
            uint8_t *o = dst + ptrdiff_t(j) * stride + i * 4;
            o[0] = static_cast<uint8_t>(t + 255.0f * linear_to_srgb(sat(c.r)));
            o[1] = static_cast<uint8_t>(t + 255.0f * linear_to_srgb(sat(c.g)));
            o[2] = static_cast<uint8_t>(t + 255.0f * linear_to_srgb(sat(c.b)));
            o[3] = static_cast<uint8_t>(t + 255.0f * sat(c.a));
        }
}

// put_image_data, hpp:3383-3408.
void oracle_write_rgba8(void *canvas, const uint8_t *src, int width, int height, int stride, int x, int y)
{
    Canvas &cv = *static_cast<Canvas *>(canvas);
    for (int j = 0; j < height; ++j)
        for (int i = 0; i < width; ++i) {
            int cx = x + i, cy = y + j;
            if (cx < 0 || cv.w <= cx || cy < 0 || cv.h <= cy) continue;
            cv.fb[size_t(cy) * size_t(cv.w) + size_t(cx)] = decode_texel(src + ptrdiff_t(j) * stride + i * 4);
        }
}

void oracle_read_f32(void *canvas, float *dst)
{
    Canvas &cv = *static_cast<Canvas *>(canvas);
    memcpy(dst, cv.fb.data(), sizeof(C) * cv.fb.size());
}

int oracle_read_mask(void *canvas, uint32_t slot, float *dst)
{
    Canvas &cv = *static_cast<Canvas *>(canvas);
    size_t n = size_t(cv.w) * size_t(cv.h);
    if (slot == 0) { for (size_t i = 0; i < n; ++i) dst[i] = 1.0f; return 0; }
    if (!cv.masks.count(slot)) return -1;
    memcpy(dst, cv.masks[slot].data(), sizeof(float) * n);
    return 0;
}

// Intermediate taps for stage-by-stage parity: flattened (+stroked) polylines of
// one draw as (x0,y0,x1,y1) closed-loop edges.  Returns the edge count.
long oracle_debug_edges(const cb200_frame *frame, uint32_t draw_index, float *edges, long capacity)
{
    const cb200_draw &d = frame->draws[draw_index];
    cb200_frame local = *frame;
    std::vector<float> pts;
    std::vector<cb200_subpath> subs;
    if (frame->n_glyphs) {                       // text drawn as glyph instances: expand them first
        expand_glyphs(frame, pts);
        subs.assign(frame->subpaths, frame->subpaths + frame->n_subpaths);
        for (size_t i = 0; i < subs.size(); ++i)
            if (subs[i].instanced) { subs[i].first_point += frame->n_points; subs[i].instanced = 0; }
        local.points = pts.data(); local.n_points = frame->n_points + frame->n_glyph_points;
        local.subpaths = subs.data(); local.n_glyphs = 0;
    }
    frame = &local;
    Poly lines, work;
    flatten(frame, d, lines);
    if (d.kind == CB200_STROKE) {
        if (d.n_dash) {
            dash(lines, frame->dashes + d.first_dash, d.n_dash, d.dash_offset, mat(d.inverse), work);
            lines.pts.swap(work.pts);
            lines.subs.swap(work.subs);
        }
        StrokeStyle st;
        st.half = d.line_width * 0.5f;
        st.miter2 = d.miter_limit * d.miter_limit * st.half * st.half;
        st.cap = d.cap; st.join = d.join;
        st.fwd = mat(d.forward); st.inv = mat(d.inverse);
        outline_stroke(lines, st, work);
        lines.pts.swap(work.pts);
        lines.subs.swap(work.subs);
    }
    long n = 0;
    size_t base = 0;
    for (size_t s = 0; s < lines.subs.size(); ++s) {
        size_t cnt = lines.subs[s].first;
        for (size_t i = 0; i < cnt; ++i, ++n) {
            if (n >= capacity) continue;
            P a = lines.pts[base + (i ? i : cnt) - 1], b = lines.pts[base + i];
            edges[n * 4 + 0] = a.x; edges[n * 4 + 1] = a.y; edges[n * 4 + 2] = b.x; edges[n * 4 + 3] = b.y;
        }
        base += cnt;
    }
    return n;
}

// Function-pointer friendly taps for cv_create_tapped (user = oracle canvas).
// is_point_in_path, hpp:3101-3132, over the already flattened path: edges are (from, to) pairs with
// every subpath's closing edge included (the reference closes them with `beginning`, hpp:3118).
void oracle_points_in_path(const float *edges, uint32_t n_edges, const float *xy, uint32_t n, uint8_t *inside)
{
    for (uint32_t i = 0; i < n; ++i) {
        const float x = xy[2 * i], y = xy[2 * i + 1];
        int winding = 0;
        bool on_edge = false;
        for (uint32_t k = 0; k < n_edges && !on_edge; ++k) {
            P from = mk(edges[4 * k], edges[4 * k + 1]), to = mk(edges[4 * k + 2], edges[4 * k + 3]);
            if ((from.y < y && y <= to.y) || (to.y < y && y <= from.y)) {
                float side = dotp(rot90(sub(to, from)), sub(mk(x, y), from));
                if (side == 0.0f) on_edge = true;
                else winding += side > 0.0f ? 1 : -1;
            } else if (from.y == y && y == to.y && ((from.x <= x && x <= to.x) || (to.x <= x && x <= from.x)))
                on_edge = true;
        }
        inside[i] = (on_edge || winding != 0) ? 1 : 0;
    }
}

// The check of the CUDA path's working-rectangle rule (canvas_ity_b200/csrc/device/edge_clip.cuh) that runs
// without a GPU: out[0..3] = the run bounding box render_shadow derives (hpp:2409-2419) from the reference's
// polygon clip (scan_convert above); out[4..7] = the join of what `per_loop` -- the product's own host build
// of its per-edge clip, scanline walk and boundary-segment walk, cb200_debug_shadow_box -- reports for each
// loop of the draw's outline.  per_loop(xy, n, off_x, off_y, padded_w, padded_h, box5): box5 = (min_x, max_x,
// min_y, max_y, first run key y << 16 | x or -1).  Boxes are (min_x, max_x, min_y, max_y); empty = all -1.
typedef void (*loop_box_fn)(const float *xy, uint32_t n, float off_x, float off_y, int pw, int ph, int *box5);

int oracle_debug_shadow_boxes(const cb200_frame *frame, uint32_t draw_index, int width, int height,
                              loop_box_fn per_loop, int *out)
{
    const cb200_draw &d = frame->draws[draw_index];
    for (int i = 0; i < 8; ++i) out[i] = -1;
    if (d.kind == CB200_CLIP) return 0;
    cb200_frame local = *frame;
    std::vector<float> pts;
    std::vector<cb200_subpath> subs;
    if (frame->n_glyphs) {
        expand_glyphs(frame, pts);
        subs.assign(frame->subpaths, frame->subpaths + frame->n_subpaths);
        for (size_t i = 0; i < subs.size(); ++i)
            if (subs[i].instanced) { subs[i].first_point += frame->n_points; subs[i].instanced = 0; }
        local.points = pts.data(); local.n_points = frame->n_points + frame->n_glyph_points;
        local.subpaths = subs.data(); local.n_glyphs = 0;
    }
    Poly lines, work;
    flatten(&local, d, lines);
    if (d.kind == CB200_STROKE) {
        if (d.n_dash) {
            dash(lines, local.dashes + d.first_dash, d.n_dash, d.dash_offset, mat(d.inverse), work);
            lines.pts.swap(work.pts); lines.subs.swap(work.subs);
        }
        StrokeStyle st;
        st.half = d.line_width * 0.5f;
        st.miter2 = d.miter_limit * d.miter_limit * st.half * st.half;
        st.cap = d.cap; st.join = d.join;
        st.fwd = mat(d.forward); st.inv = mat(d.inverse);
        outline_stroke(lines, st, work);
        lines.pts.swap(work.pts); lines.subs.swap(work.subs);
    }
    float sigma2 = 0.25f * d.shadow_blur * d.shadow_blur;
    size_t radius = size_t(0.5f * sqrtf(4.0f * sigma2 + 1.0f) - 0.5f);
    int border = 3 * (int(radius) + 1);
    P off = mk(float(border) + d.shadow_offset_x, float(border) + d.shadow_offset_y);
    const int pw = width + 2 * border, ph = height + 2 * border;
    Coverage cov;
    scan_convert(lines, off, pw, ph, cov);
    if (!cov.empty()) { out[0] = cov.min_x; out[1] = cov.max_x; out[2] = cov.min_y; out[3] = cov.max_y; }
    int lx = 1 << 30, hx = -1, ly = 1 << 30, hy = -1;
    long first_key = -1;
    auto enter = [&](int x, int y) { lx = std::min(lx, x); hx = std::max(hx, x); ly = std::min(ly, y); hy = std::max(hy, y); };
    size_t base = 0;
    for (size_t s = 0; s < lines.subs.size(); ++s) {
        size_t n = lines.subs[s].first;
        if (per_loop && n) {
            int box[5] = { 0, -1, 0, -1, -1 };
            per_loop(&lines.pts[base].x, uint32_t(n), off.x, off.y, pw, ph, box);
            if (box[1] >= 0 && box[3] >= 0) { enter(box[0], box[2]); enter(box[1], box[3]); }
            if (box[4] >= 0 && (first_key < 0 || box[4] < first_key)) first_key = box[4];
        }
        base += n;
    }
    if (first_key >= 0) enter(int(first_key & 0xffff), int(first_key >> 16));
    if (hx >= 0 && hy >= 0) { out[4] = lx; out[5] = hx; out[6] = ly; out[7] = hy; }
    return border;
}

// Stage tap for tests/test_scan_conversion.py: the outline of one draw (flattened, dashed, stroked) as closed loops
// -- counts[l] points each, xy packed -- so a test can hand single loops to both scan converters.  Returns the number
// of loops; *n_points the total point count (copies at most the capacities).
long oracle_debug_loops(const cb200_frame *frame, uint32_t draw_index, float *xy, long cap_points, uint32_t *counts,
                        long cap_loops, long *n_points)
{
    const cb200_draw &d = frame->draws[draw_index];
    cb200_frame local = *frame;
    std::vector<float> pts;
    std::vector<cb200_subpath> subs;
    if (frame->n_glyphs) {
        expand_glyphs(frame, pts);
        subs.assign(frame->subpaths, frame->subpaths + frame->n_subpaths);
        for (size_t i = 0; i < subs.size(); ++i)
            if (subs[i].instanced) { subs[i].first_point += frame->n_points; subs[i].instanced = 0; }
        local.points = pts.data(); local.n_points = frame->n_points + frame->n_glyph_points;
        local.subpaths = subs.data(); local.n_glyphs = 0;
    }
    Poly lines, work;
    flatten(&local, d, lines);
    if (d.kind == CB200_STROKE) {
        if (d.n_dash) {
            dash(lines, local.dashes + d.first_dash, d.n_dash, d.dash_offset, mat(d.inverse), work);
            lines.pts.swap(work.pts); lines.subs.swap(work.subs);
        }
        StrokeStyle st;
        st.half = d.line_width * 0.5f;
        st.miter2 = d.miter_limit * d.miter_limit * st.half * st.half;
        st.cap = d.cap; st.join = d.join;
        st.fwd = mat(d.forward); st.inv = mat(d.inverse);
        outline_stroke(lines, st, work);
        lines.pts.swap(work.pts); lines.subs.swap(work.subs);
    }
    for (size_t i = 0; i < lines.pts.size() && long(i) < cap_points; ++i) { xy[2 * i] = lines.pts[i].x; xy[2 * i + 1] = lines.pts[i].y; }
    for (size_t s = 0; s < lines.subs.size() && long(s) < cap_loops; ++s) counts[s] = uint32_t(lines.subs[s].first);
    if (n_points) *n_points = long(lines.pts.size());
    return long(lines.subs.size());
}

// The reference's scan conversion (polygon clip + add_runs, hpp:2193-2240, before the sort) of ONE closed loop:
// every run as (x, y, delta).  Returns the count, copies at most `capacity`.
long oracle_debug_loop_runs(const float *xy, uint32_t n, float off_x, float off_y, int padded_w, int padded_h,
                            int32_t *run_xy, float *run_delta, long capacity)
{
    Poly one;
    for (uint32_t i = 0; i < n; ++i) one.pts.push_back(mk(xy[2 * i], xy[2 * i + 1]));
    one.subs.push_back(std::make_pair(size_t(n), true));
    Coverage cov;
    scan_convert_raw(one, mk(off_x, off_y), padded_w, padded_h, cov);
    long count = 0;
    for (int y = 0; y < cov.rows; ++y)
        for (size_t r = 0; r < cov.row[size_t(y)].size(); ++r, ++count)
            if (count < capacity) {
                if (run_xy) { run_xy[2 * count] = int32_t(cov.row[size_t(y)][r].x); run_xy[2 * count + 1] = y; }
                if (run_delta) run_delta[count] = cov.row[size_t(y)][r].d;
            }
    return count;
}

void oracle_tap_frame(void *user, const cb200_frame *frame) { oracle_submit(user, frame); }
void oracle_tap_read(void *user, uint8_t *dst, int w, int h, int stride, int x, int y)
{
    oracle_read_rgba8(user, dst, w, h, stride, x, y);
}
void oracle_tap_write(void *user, const uint8_t *src, int w, int h, int stride, int x, int y)
{
    oracle_write_rgba8(user, src, w, h, stride, x, y);
}

}  // extern "C"
