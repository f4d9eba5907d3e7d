// geom.cuh -- curve flattening shared by the flatten kernel (device) and the
// host front end (is_point_in_path needs a synchronous flattening, reference
// src/canvas_ity.hpp:3101-3132).
//
// Behavioural spec (reference "hpp" = src/canvas_ity.hpp):
//   flatten_cubic      == add_bezier        hpp:1398-1487
//   subdivide_piece    == add_tessellation  hpp:1331-1387 (recursion -> explicit stack)
//   stroke_angular     == path_to_lines     hpp:1498-1500
// Every float operation keeps the reference's association order and this file
// must be compiled without FMA contraction (-fmad=false / -ffp-contract=off):
// the flatness, angle and cut tests are discontinuous decisions.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb200 {

struct vec2 { float x, y; };

CB_HD vec2 v2(float x, float y) { vec2 r; r.x = x; r.y = y; return r; }
CB_HD vec2 operator+(vec2 a, vec2 b) { return v2(a.x + b.x, a.y + b.y); }
CB_HD vec2 operator-(vec2 a, vec2 b) { return v2(a.x - b.x, a.y - b.y); }
CB_HD vec2 operator*(float s, vec2 a) { return v2(a.x * s, a.y * s); }
CB_HD float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
CB_HD vec2 perp(vec2 a) { return v2(-a.y, a.x); }
CB_HD vec2 mix(vec2 a, vec2 b, float t) { return a + t * (b - a); }
CB_HD float vlen(vec2 a) { return sqrtf(dot(a, a)); }
CB_HD vec2 unit(vec2 a) { return (1.0f / fmaxf(1.0e-6f, vlen(a))) * a; }
CB_HD float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

struct affine { float a, b, c, d, e, f; };
CB_HD vec2 apply(const affine &m, vec2 p) {
    return v2(m.a * p.x + m.c * p.y + m.e, m.b * p.x + m.d * p.y + m.f);
}

// ---- acosf / tanf of the rounded join (hpp:1995-1997), bit for bit --------------------------------
// The reference calls the host libm there, and the join's cubic then goes through flattening DECISIONS, so a
// last-bit difference in the angle can change the outline.  glibc (2.39 on this image) implements both with the
// fdlibm algorithms (flt-32/e_acosf.c, k_tanf.c): plain float polynomial arithmetic, restated here operation by
// operation (no FMA).  Checked against glibc's acosf over ALL floats in [-1, 1] and tanf over all floats in
// [0, pi/4] -- the join only ever needs tanf(angle / 4) with angle in [0, pi]: zero mismatches
// (tests/test_geometry_math.py samples the same through cb200_debug_join_math).
CB_HD float bits_float(unsigned u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { unsigned u; float f; } c; c.u = u; return c.f;
#endif
}
CB_HD unsigned float_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { unsigned u; float f; } c; c.f = f; return c.u;
#endif
}

CB_HD float join_acosf(float x)
{
    const float pi = bits_float(0x40490fdau), pio2_hi = bits_float(0x3fc90fdau), pio2_lo = bits_float(0x33a22168u);
    const float p0 = bits_float(0x3e2aaaabu), p1 = bits_float(0xbea6b090u), p2 = bits_float(0x3e4e0aa8u),
                p3 = bits_float(0xbd241146u), p4 = bits_float(0x3a4f7f04u), p5 = bits_float(0x3811ef08u);
    const float q1 = bits_float(0xc019d139u), q2 = bits_float(0x4001572du), q3 = bits_float(0xbf303361u),
                q4 = bits_float(0x3d9dc62eu);
    const int hx = int(float_bits(x)), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    if (ix < 0x3f000000) {                                   // |x| < 0.5
        if (ix <= 0x32800000) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = z * (p0 + z * (p1 + z * (p2 + z * (p3 + z * (p4 + z * p5)))));
        const float q = 1.0f + z * (q1 + z * (q2 + z * (q3 + z * q4)));
        const float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (hx < 0) {                                            // x < -0.5
        const float z = (1.0f + x) * 0.5f;
        const float p = z * (p0 + z * (p1 + z * (p2 + z * (p3 + z * (p4 + z * p5)))));
        const float q = 1.0f + z * (q1 + z * (q2 + z * (q3 + z * q4)));
        const float s = sqrtf(z);
        const float r = p / q;
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float z = (1.0f - x) * 0.5f;                       // x > 0.5
    const float s = sqrtf(z);
    const float df = bits_float(float_bits(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float p = z * (p0 + z * (p1 + z * (p2 + z * (p3 + z * (p4 + z * p5)))));
    const float q = 1.0f + z * (q1 + z * (q2 + z * (q3 + z * q4)));
    const float r = p / q;
    const float w = r * s + c;
    return 2.0f * (df + w);
}

// tanf(x) for 0 <= x <= pi/4 (glibc: __kernel_tanf(x, 0, 1), no argument reduction in that range)
CB_HD float join_tanf(float x)
{
    const float pio4 = bits_float(0x3f490fdau), pio4lo = bits_float(0x33222168u);
    const float t0 = bits_float(0x3eaaaaabu), t1 = bits_float(0x3e088889u), t2 = bits_float(0x3d5d0dd1u),
                t3 = bits_float(0x3cb327a4u), t4 = bits_float(0x3c11371fu), t5 = bits_float(0x3b6b6916u),
                t6 = bits_float(0x3abede48u), t7 = bits_float(0x3a1a26c8u), t8 = bits_float(0x398137b9u),
                t9 = bits_float(0x38a3f445u), t10 = bits_float(0x3895c07au), t11 = bits_float(0xb79bae5fu),
                t12 = bits_float(0x37d95384u);
    const int hx = int(float_bits(x)), ix = hx & 0x7fffffff;
    if (ix < 0x39000000 && int(x) == 0) return x;            // |x| < 2^-13
    const bool big = ix >= 0x3f2ca140;                       // |x| >= 0.6744
    if (big) {
        const float z = pio4 - x;
        x = z + pio4lo;
        if (fabsf(x) < 1.220703125e-4f) return 1.0f - 2.0f * x;    // 2^-13
    }
    const float z = x * x;
    float w = z * z;
    float r = t1 + w * (t3 + w * (t5 + w * (t7 + w * (t9 + w * t11))));
    const float v = z * (t2 + w * (t4 + w * (t6 + w * (t8 + w * (t10 + w * t12)))));
    const float s = z * x;
    r = 0.0f + z * (s * (r + v) + 0.0f);
    r += t0 * s;
    w = x + r;
    if (big) return 1.0f - 2.0f * (x - (w * w / (w + 1.0f) - r));
    return w;
}

// Cosine of the largest turn a stroked curve piece may keep (hpp:1498-1500);
// fills use -1 which also switches the emission rule (end points only).
CB_HD float stroke_angular(float line_width) {
    const float tolerance = 0.125f;
    float ratio = tolerance / fmaxf(0.5f * line_width, tolerance);
    return (ratio - 2.0f) * ratio * 2.0f + 1.0f;
}

// Flatness / turn test of one curve piece (hpp:1341-1369).  Returns true when the
// piece needs no further halving; q1..q3 are the squared control-polygon edges the
// emission rule needs.
CB_HD bool piece_is_flat(vec2 p0, vec2 c1, vec2 c2, vec2 p3, float angular, float &q1, float &q2, float &q3)
{
    vec2 e1 = c1 - p0, e2 = c2 - c1, e3 = p3 - c2, chord = p3 - p0;
    q1 = dot(e1, e1); q2 = dot(e2, e2); q3 = dot(e3, e3);
    float chord2 = fmaxf(1.0e-4f, dot(chord, chord));
    float t1 = clamp01(dot(e1, chord) / chord2);
    float t2 = clamp01(dot(e3, chord) / chord2);
    vec2 off1 = p0 + t1 * chord - c1;
    vec2 off2 = p3 - t2 * chord - c2;
    float cosine = 1.0f;
    if (angular > -1.0f) {
        if (q1 * q3 != 0.0f) cosine = dot(e1, e3) / sqrtf(q1 * q3);
        else if (q1 * q2 != 0.0f) cosine = dot(e1, e2) / sqrtf(q1 * q2);
        else if (q2 * q3 != 0.0f) cosine = dot(e2, e3) / sqrtf(q2 * q3);
    }
    const float flat2 = 0.125f * 0.125f;
    return dot(off1, off1) <= flat2 && dot(off2, off2) <= flat2 && cosine >= angular;
}

// What a finished piece contributes (hpp:1371-1376): strokes also get the control
// points so that joins see true tangents; fills only the end point.
template <class Sink>
CB_HD void emit_flat_piece(vec2 c1, vec2 c2, vec2 p3, float angular, float q1, float q2, float q3, Sink &sink)
{
    const bool stroking = angular > -1.0f;
    if (stroking && q1 != 0.0f) sink.put(c1);
    if (stroking && q2 != 0.0f) sink.put(c2);
    if (angular == -1.0f || q3 != 0.0f) sink.put(p3);
}

// One monotone piece: halve until both inner control points sit within 1/8 px
// of the chord and (strokes) the turn stays under the angular limit; depth budget
// 20 from the top.  Left halves first so points come out in curve order.
// `Sink::put(vec2)`.
template <class Sink>
CB_HD void subdivide_piece(vec2 p0, vec2 c1, vec2 c2, vec2 p3, float angular, Sink &sink, int budget = 20)
{
    // pending right halves; their start point is wherever the left subtree ends
    vec2 stack_c1[21], stack_c2[21], stack_p3[21];
    int stack_budget[21];
    int depth = 0;
    for (;;) {
        float q1, q2, q3;
        bool done = piece_is_flat(p0, c1, c2, p3, angular, q1, q2, q3) || budget == 0;
        if (done) {
            emit_flat_piece(c1, c2, p3, angular, q1, q2, q3, sink);
            if (depth == 0) return;
            --depth;
            p0 = p3;                       // right half starts where we stopped
            c1 = stack_c1[depth]; c2 = stack_c2[depth]; p3 = stack_p3[depth];
            budget = stack_budget[depth];
            continue;
        }
        vec2 l1 = mix(p0, c1, 0.5f);
        vec2 mid = mix(c1, c2, 0.5f);
        vec2 r2 = mix(c2, p3, 0.5f);
        vec2 l2 = mix(l1, mid, 0.5f);
        vec2 r1 = mix(mid, r2, 0.5f);
        vec2 split = mix(l2, r1, 0.5f);
        --budget;
        stack_c1[depth] = r1; stack_c2[depth] = r2; stack_p3[depth] = p3;
        stack_budget[depth] = budget;
        ++depth;
        c1 = l1; c2 = l2; p3 = split;
    }
}

// Roots of the derivative's quadratic a t^2 + b t + c on one axis, in the
// cancellation-free form (hpp:1420-1434).  Appends to cut[n...].
CB_HD int axis_extrema(float a, float b, float c, float *cut, int n)
{
    const float eps = 1.0e-4f;
    if (fabsf(a) > eps) {
        float disc = b * b - 4.0f * a * c;
        if (disc >= 0.0f) {
            float sgn = b > 0.0f ? 1.0f : -1.0f;
            float q = -b - sgn * sqrtf(disc);
            float r = q / (2.0f * a);
            cut[n++] = r;
            cut[n++] = c / (a * r);
        }
    } else if (fabsf(b) > eps)
        cut[n++] = -c / b;
    return n;
}

// Cut parameters of a cubic (hpp:1414-1465): 0, 1, the x/y extrema and the curvature
// extremum, insertion-sorted.  Returns how many.
CB_HD int cubic_cuts(vec2 p0, vec2 c1, vec2 c2, vec2 p3, float *cut)
{
    vec2 e1 = c1 - p0, e2 = c2 - c1, e3 = p3 - c2;
    cut[0] = 0.0f; cut[1] = 1.0f;
    int n = 2;
    vec2 qa = -9.0f * e2 + 3.0f * (p3 - p0);
    vec2 qb = 6.0f * (p0 + c2) - 12.0f * c1;
    vec2 qc = 3.0f * e1;
    n = axis_extrema(qa.x, qb.x, qc.x, cut, n);
    n = axis_extrema(qa.y, qb.y, qc.y, cut, n);
    float d12 = dot(perp(e1), e2), d13 = dot(perp(e1), e3), d23 = dot(perp(e2), e3);
    float ka = d12 - d13 + d23;
    float kb = -2.0f * d12 + d13;
    if (fabsf(ka) > 1.0e-4f && fabsf(kb) > 1.0e-4f)
        cut[n++] = -0.5f * kb / ka;
    for (int i = 1; i < n; ++i) {          // tiny insertion sort, NaNs stay put
        float v = cut[i];
        int j = i - 1;
        for (; j >= 0 && v < cut[j]; --j) cut[j + 1] = cut[j];
        cut[j + 1] = v;
    }
    return n;
}

CB_HD bool cut_is_kept(float lo, float hi) { return 0.0f <= lo && hi <= 1.0f && lo != hi; }

// Control points of the piece [lo, hi] by blossoming (hpp:1472-1481): de Casteljau
// at hi, then again at lo/hi from the left.  `from` is where the previous piece ended.
CB_HD void cubic_piece(vec2 p0, vec2 c1, vec2 c2, vec2 p3, float lo, float hi, vec2 &k1, vec2 &k2, vec2 &to)
{
    float rel = lo / hi;
    vec2 a1 = mix(p0, c1, hi), a2 = mix(c1, c2, hi), a3 = mix(c2, p3, hi);
    vec2 b1 = mix(a1, a2, hi), b2 = mix(a2, a3, hi);
    vec2 g = mix(a1, b1, rel);
    to = mix(b1, b2, hi);
    k2 = mix(b1, to, rel);
    k1 = mix(g, k2, rel);
}

// A whole cubic: cut at x/y extrema and at the curvature extremum so every
// piece is monotone and turns < 90 degrees, then subdivide each piece.
template <class Sink>
CB_HD void flatten_cubic(vec2 p0, vec2 c1, vec2 c2, vec2 p3, float angular, Sink &sink)
{
    vec2 e1 = c1 - p0, e3 = p3 - c2;
    if (dot(e1, e1) == 0.0f && dot(e3, e3) == 0.0f) {   // a line (or a point)
        sink.put(p3);
        return;
    }
    float cut[7];
    int n = cubic_cuts(p0, c1, c2, p3, cut);
    vec2 from = p0;
    for (int i = 0; i + 1 < n; ++i) {
        if (!cut_is_kept(cut[i], cut[i + 1])) continue;
        vec2 k1, k2, to;
        cubic_piece(p0, c1, c2, p3, cut[i], cut[i + 1], k1, k2, to);
        subdivide_piece(from, k1, k2, to, angular, sink);
        from = to;
    }
}

struct count_sink { int n; CB_HD void put(vec2) { ++n; } };
struct write_sink { vec2 *out; int n; CB_HD void put(vec2 p) { out[n++] = p; } };

} // namespace cb200
