// composite.cu -- K7: tile compositor.  One CTA owns one 32x32 framebuffer tile,
// split into four independent warps of 8 scanlines; each keeps its 256 linear
// premultiplied float RGBA pixels in registers (8 per lane), replays every job
// that touches the tile IN SUBMISSION ORDER and stores its pixels once.  Per job and pixel it does what the reference does per span
// pixel: coverage from the sorted runs (tile_cov.cuh), paint_pixel (hpp:2265-2377),
// the 4-bit Porter-Duff mix and the visibility lerp (hpp:2570-2591).  Shadow jobs
// take their coverage from the blurred plane instead (hpp:2504-2538) and clip jobs
// write coverage * visibility into a new mask plane (hpp:3057-3099).
//
// Framebuffer traffic is one 16-byte vector load and one 16-byte vector store per
// pixel per FRAME (512 B contiguous per warp row), however many draws overlap --
// the per-draw 32 B/pixel of the reference's read-modify-write loop stays in
// registers.
#include "frame.cuh"
#include "tile_cov.cuh"

#include <cstdlib>

namespace cb200 {

namespace {

struct rgba { float r, g, b, a; };
__device__ __forceinline__ rgba mk(float r, float g, float b, float a) { rgba c = { r, g, b, a }; return c; }
__device__ __forceinline__ rgba scale(float s, rgba c) { return mk(c.r * s, c.g * s, c.b * s, c.a * s); }
__device__ __forceinline__ rgba plus(rgba x, rgba y) { return mk(x.r + y.r, x.g + y.g, x.b + y.b, x.a + y.a); }

__device__ __forceinline__ float keys_weight(float t)
{
    return t < 1.0f ? (1.5f * t - 2.5f) * t * t + 1.0f : ((-0.5f * t + 2.5f) * t - 4.0f) * t + 2.0f;
}

struct paint_tables {
    const float4 *colors; const float *stops; const float4 *texels;
    const brush_rec *brushes; const draw_rec *draws;
};

// paint_pixel, hpp:2265-2377.  `at` is the device-space pixel centre.
__device__ rgba paint_at(const paint_tables &f, const brush_rec &b, const affine &inv, vec2 at)
{
    if (b.n_colors == 0) return mk(0.0f, 0.0f, 0.0f, 0.0f);
    if (b.type == CB200_BRUSH_COLOR) {
        float4 c = f.colors[b.first_color];
        return mk(c.x, c.y, c.z, c.w);
    }
    vec2 p = apply(inv, at);
    if (b.type == CB200_BRUSH_PATTERN) {
        float w = float(b.width), h = float(b.height);
        if (((b.repetition & 2u) && (p.x < 0.0f || w <= p.x)) ||
            ((b.repetition & 1u) && (p.y < 0.0f || h <= p.y)))
            return mk(0.0f, 0.0f, 0.0f, 0.0f);
        float sx = fabsf(inv.a) + fabsf(inv.c), sy = fabsf(inv.b) + fabsf(inv.d);
        sx = fmaxf(1.0f, fminf(sx, w * 0.25f));
        sy = fmaxf(1.0f, fminf(sy, h * 0.25f));
        float rx = 1.0f / sx, ry = 1.0f / sy;
        p = p - v2(0.5f, 0.5f);
        int x0 = int(ceilf(p.x - sx * 2.0f)), y0 = int(ceilf(p.y - sy * 2.0f));
        int x1 = int(ceilf(p.x + sx * 2.0f)), y1 = int(ceilf(p.y + sy * 2.0f));
        const float4 *tex = f.texels + b.texel_offset;
        const bool clamp_mode = (b.flags & CB200_BRUSH_CLAMP) != 0;
        rgba acc = mk(0.0f, 0.0f, 0.0f, 0.0f);
        float wsum = 0.0f;
        for (int ty = y0; ty < y1; ++ty) {
            float wy = keys_weight(fabsf(ry * (float(ty) - p.y)));
            int yy = ty % b.height;
            if (yy < 0) yy += b.height;
            if (clamp_mode) yy = min(max(ty, 0), b.height - 1);
            const float4 *row = tex + size_t(yy) * size_t(b.width);
            for (int tx = x0; tx < x1; ++tx) {
                float wx = keys_weight(fabsf(rx * (float(tx) - p.x)));
                int xx = tx % b.width;
                if (xx < 0) xx += b.width;
                if (clamp_mode) xx = min(max(tx, 0), b.width - 1);
                float wgt = wx * wy;
                float4 c = __ldg(row + xx);
                acc = plus(acc, scale(wgt, mk(c.x, c.y, c.z, c.w)));
                wsum += wgt;
            }
        }
        return scale(1.0f / wsum, acc);
    }
    vec2 rel = p - v2(b.sx, b.sy), axis = v2(b.ex, b.ey) - v2(b.sx, b.sy);
    float along = dot(rel, axis), axis2 = dot(axis, axis);
    float t;
    if (b.type == CB200_BRUSH_LINEAR) {
        if (axis2 == 0.0f) return mk(0.0f, 0.0f, 0.0f, 0.0f);
        t = along / axis2;
    } else {
        float dr = b.r1 - b.r0;
        float qa = axis2 - dr * dr;
        float qb = -2.0f * (along + b.r0 * dr);
        float qc = dot(rel, rel) - b.r0 * b.r0;
        float disc = qb * qb - 4.0f * qa * qc;
        if (disc < 0.0f || (axis2 == 0.0f && dr == 0.0f)) return mk(0.0f, 0.0f, 0.0f, 0.0f);
        float root = sqrtf(disc), inv2a = 1.0f / (2.0f * qa);
        float ta = (-qb - root) * inv2a, tb = (-qb + root) * inv2a;
        if (b.r0 + dr * tb >= 0.0f) t = tb;
        else if (b.r0 + dr * ta >= 0.0f) t = ta;
        else return mk(0.0f, 0.0f, 0.0f, 0.0f);
    }
    // first stop strictly greater than t (upper_bound): NaN compares false everywhere
    const float *stops = f.stops + b.first_color;
    uint32_t hi = 0;
    while (hi < b.n_colors && !(t < stops[hi])) ++hi;
    float4 c;
    if (hi == 0) c = f.colors[b.first_color];
    else if (hi == b.n_colors) c = f.colors[b.first_color + b.n_colors - 1];
    else {
        float m = (t - stops[hi - 1]) / (stops[hi] - stops[hi - 1]);
        float4 lo = f.colors[b.first_color + hi - 1], up = f.colors[b.first_color + hi];
        c = make_float4(lo.x + m * (up.x - lo.x), lo.y + m * (up.y - lo.y), lo.z + m * (up.z - lo.z),
                        lo.w + m * (up.w - lo.w));
    }
    return mk(c.x * c.w, c.y * c.w, c.z * c.w, c.w);
}

// the mix program of hpp:2583-2591.  The products are fused (fmaf): the reference's own builds
// differ by more than that between compilers, and with cov = vis = 1 and an opaque
// source the result is still exactly `fore`, which is what occlusion culling relies on.
__device__ __forceinline__ void blend(float4 &back, rgba fore, uint32_t op, float vis)
{
    float mf = (op & 1u) ? back.w : 0.0f;
    if (op & 2u) mf = 1.0f - mf;
    float mb = (op & 4u) ? fore.a : 0.0f;
    if (op & 8u) mb = 1.0f - mb;
    float r = fmaf(mf, fore.r, mb * back.x), g = fmaf(mf, fore.g, mb * back.y), b = fmaf(mf, fore.b, mb * back.z);
    float a = fminf(fmaf(mf, fore.a, mb * back.w), 1.0f);
    float keep = 1.0f - vis;
    back = make_float4(fmaf(vis, r, keep * back.x), fmaf(vis, g, keep * back.y), fmaf(vis, b, keep * back.z),
                       fmaf(vis, a, keep * back.w));
}

// The same program where nothing clips the draw (visibility exactly 1): the closing lerp
// back = 1 * blend + 0 * back is the blend itself, and source_over (mix_fore = 1, mix_back = 1 - fore.a,
// three draws out of four) needs no selects.  Same products and sums as blend(), so the same bits
// (for finite pixels; 0 * inf would be NaN in the lerp).
__device__ __forceinline__ void blend_unclipped(float4 &back, rgba fore, uint32_t op)
{
    if (op == 14u) {
        const float mb = 1.0f - fore.a;
        back = make_float4(fore.r + mb * back.x, fore.g + mb * back.y, fore.b + mb * back.z,
                           fminf(fore.a + mb * back.w, 1.0f));
        return;
    }
    float mf = (op & 1u) ? back.w : 0.0f;
    if (op & 2u) mf = 1.0f - mf;
    float mb = (op & 4u) ? fore.a : 0.0f;
    if (op & 8u) mb = 1.0f - mb;
    back = make_float4(fmaf(mf, fore.r, mb * back.x), fmaf(mf, fore.g, mb * back.y), fmaf(mf, fore.b, mb * back.z),
                       fminf(fmaf(mf, fore.a, mb * back.w), 1.0f));
}

constexpr int kWarpRows = 8;                                 // scanlines of a tile owned by one warp
constexpr int kTileWarps = kTile / kWarpRows;               // 4 warps per tile
constexpr int kCompBlock = 32 * kTileWarps;                 // one CTA = one tile = 128 threads

// A non-solid brush staged in shared memory once per (job, warp): record, brush-space matrix and
// up to kStagedStops gradient stops, so the per-pixel code touches no global tables.
constexpr uint32_t kStagedStops = 16;
struct staged_brush {
    brush_rec b;
    affine inv;
    float stops[kStagedStops];
    float inv_span[kStagedStops];          // 1 / (stops[k] - stops[k-1]), k >= 1
    float4 colors[kStagedStops];
    // pattern fills under an axis-aligned matrix: Keys weights, texel-row offsets and flags of the warp's 8 scanlines
    float4 row_w[8];
    int4 row_i[8];
    int row_flags[8];
};

// Linear / radial gradient at a device-space pixel centre (hpp:2331-2376), from the staged copy.
// Everything that does not depend on the pixel is gathered once per (job, warp): the axis, the radial
// quadratic's leading coefficient and its reciprocal, and -- for the usual handful of stops -- the stop
// positions themselves, so that the upper_bound search is four chained compares instead of a loop over
// shared memory.  The offset inside a stop interval is multiplied by the interval's reciprocal (staged)
// where the reference divides: at most 2 ulp of the interpolation factor, no decision depends on it.
struct gradient_ctx {
    affine inv;
    float sx, sy, ax, ay, axis2, r0, dr, qa, inv2a;
    float s0, s1, s2, s3;
    uint32_t n;
    bool linear;
};

__device__ __forceinline__ gradient_ctx make_gradient_ctx(const staged_brush &sb)
{
    gradient_ctx g;
    const brush_rec &b = sb.b;
    g.inv = sb.inv;
    g.sx = b.sx; g.sy = b.sy;
    vec2 axis = v2(b.ex, b.ey) - v2(b.sx, b.sy);
    g.ax = axis.x; g.ay = axis.y;
    g.axis2 = dot(axis, axis);
    g.r0 = b.r0; g.dr = b.r1 - b.r0;
    g.qa = g.axis2 - g.dr * g.dr;
    g.inv2a = 1.0f / (2.0f * g.qa);
    g.n = b.n_colors;
    g.linear = b.type == CB200_BRUSH_LINEAR;
    const float never = __int_as_float(0x7f800000);        // t < +inf: the search stops there
    g.s0 = g.n > 0 ? sb.stops[0] : never; g.s1 = g.n > 1 ? sb.stops[1] : never;
    g.s2 = g.n > 2 ? sb.stops[2] : never; g.s3 = g.n > 3 ? sb.stops[3] : never;
    return g;
}

__device__ __forceinline__ rgba gradient_at(const gradient_ctx &g, const staged_brush &sb, float x, float y)
{
    vec2 p = apply(g.inv, v2(x, y));
    vec2 rel = p - v2(g.sx, g.sy), axis = v2(g.ax, g.ay);
    float along = dot(rel, axis);
    float t;
    if (g.linear) {
        if (g.axis2 == 0.0f) return mk(0.0f, 0.0f, 0.0f, 0.0f);
        t = along / g.axis2;
    } else {
        float qb = -2.0f * (along + g.r0 * g.dr);
        float qc = dot(rel, rel) - g.r0 * g.r0;
        float disc = qb * qb - 4.0f * g.qa * qc;
        if (disc < 0.0f || (g.axis2 == 0.0f && g.dr == 0.0f)) return mk(0.0f, 0.0f, 0.0f, 0.0f);
        float root = sqrtf(disc);
        float ta = (-qb - root) * g.inv2a, tb = (-qb + root) * g.inv2a;
        if (g.r0 + g.dr * tb >= 0.0f) t = tb;
        else if (g.r0 + g.dr * ta >= 0.0f) t = ta;
        else return mk(0.0f, 0.0f, 0.0f, 0.0f);
    }
    uint32_t hi = 0;                                    // first stop strictly greater than t (NaN: none)
    if (g.n <= 4) {
        const bool c0 = !(t < g.s0), c1 = c0 && !(t < g.s1), c2 = c1 && !(t < g.s2), c3 = c2 && !(t < g.s3);
        hi = min(uint32_t(c0) + uint32_t(c1) + uint32_t(c2) + uint32_t(c3), g.n);
    } else {
        while (hi < g.n && !(t < sb.stops[hi])) ++hi;
    }
    float4 c;
    if (hi == 0) c = sb.colors[0];
    else if (hi == g.n) c = sb.colors[g.n - 1];
    else {
        float m = (t - sb.stops[hi - 1]) * sb.inv_span[hi];
        float4 lo = sb.colors[hi - 1], up = sb.colors[hi];
        c = make_float4(fmaf(m, up.x - lo.x, lo.x), fmaf(m, up.y - lo.y, lo.y), fmaf(m, up.z - lo.z, lo.z),
                        fmaf(m, up.w - lo.w, lo.w));
    }
    return mk(c.x * c.w, c.y * c.w, c.z * c.w, c.w);
}

__device__ __noinline__ rgba paint_gradient(const staged_brush &sb, float x, float y)
{
    return gradient_at(make_gradient_ctx(sb), sb, x, y);
}

// Bicubic (Keys) pattern / image sample (hpp:2274-2330).  Same taps, weights and summation order as
// the reference (rows outer, columns inner), but the column weights and wrapped column indices are
// computed once per pixel instead of once per tap, and clamp addressing skips the modulo.
constexpr int kMaxTapsX = 8;
__device__ __noinline__ rgba paint_pattern(const float4 *texels, const staged_brush *sbp, float x, float y)
{
    const staged_brush &sb = *sbp;
    const brush_rec &b = sb.b;
    const affine &inv = sb.inv;
    vec2 p = apply(inv, v2(x, y));
    float w = float(b.width), h = float(b.height);
    if (((b.repetition & 2u) && (p.x < 0.0f || w <= p.x)) || ((b.repetition & 1u) && (p.y < 0.0f || h <= p.y)))
        return mk(0.0f, 0.0f, 0.0f, 0.0f);
    float sx = fabsf(inv.a) + fabsf(inv.c), sy = fabsf(inv.b) + fabsf(inv.d);
    sx = fmaxf(1.0f, fminf(sx, w * 0.25f));
    sy = fmaxf(1.0f, fminf(sy, h * 0.25f));
    float rx = 1.0f / sx, ry = 1.0f / sy;
    p = p - v2(0.5f, 0.5f);
    int x0 = int(ceilf(p.x - sx * 2.0f)), y0 = int(ceilf(p.y - sy * 2.0f));
    int x1 = int(ceilf(p.x + sx * 2.0f)), y1 = int(ceilf(p.y + sy * 2.0f));
    const float4 *tex = texels + b.texel_offset;
    const bool clamp_mode = (b.flags & CB200_BRUSH_CLAMP) != 0;
    auto wrap = [clamp_mode](int i, int n) -> int {
        if (clamp_mode) return min(max(i, 0), n - 1);
        int m = i % n;
        return m < 0 ? m + n : m;
    };
    rgba acc = mk(0.0f, 0.0f, 0.0f, 0.0f);
    float wsum = 0.0f;
    const int nx = x1 - x0;
    if (nx <= kMaxTapsX) {
        float wx[kMaxTapsX];
        int ix[kMaxTapsX];
#pragma unroll
        for (int k = 0; k < kMaxTapsX; ++k) {
            wx[k] = k < nx ? keys_weight(fabsf(rx * (float(x0 + k) - p.x))) : 0.0f;
            ix[k] = k < nx ? wrap(x0 + k, b.width) : 0;
        }
        for (int ty = y0; ty < y1; ++ty) {
            float wy = keys_weight(fabsf(ry * (float(ty) - p.y)));
            const float4 *row = tex + size_t(wrap(ty, b.height)) * size_t(b.width);
#pragma unroll
            for (int k = 0; k < kMaxTapsX; ++k) {
                if (k < nx) {
                    float wgt = wx[k] * wy;
                    float4 c = __ldg(row + ix[k]);
                    acc = plus(acc, scale(wgt, mk(c.x, c.y, c.z, c.w)));
                    wsum += wgt;
                }
            }
        }
    } else {
        for (int ty = y0; ty < y1; ++ty) {
            float wy = keys_weight(fabsf(ry * (float(ty) - p.y)));
            const float4 *row = tex + size_t(wrap(ty, b.height)) * size_t(b.width);
            for (int tx = x0; tx < x1; ++tx) {
                float wgt = keys_weight(fabsf(rx * (float(tx) - p.x))) * wy;
                float4 c = __ldg(row + wrap(tx, b.width));
                acc = plus(acc, scale(wgt, mk(c.x, c.y, c.z, c.w)));
                wsum += wgt;
            }
        }
    }
    return scale(1.0f / wsum, acc);
}

// The same sample for the common footprint of four taps per axis (no minification: scale clamps to 1),
// with the work that is shared taken out of the pixel: the taps of one axis -- first index, Keys weights,
// wrapped texel indices (one modulo, then increments) -- depend on one brush-space coordinate only, so an
// axis-aligned brush matrix gives every row of a lane the same column taps and every lane of a row the same
// row taps.  Same taps, weights (weight_x * weight_y) and summation order as hpp:2296-2328; the products are
// fused into the sums.
struct axis_taps { float w[4]; int i[4]; int n; };
__device__ __forceinline__ axis_taps keys_taps(float q, float s, float rcp, int size, bool clamp_mode)
{
    axis_taps a;
    const int first = int(ceilf(q - s * 2.0f)), last = int(ceilf(q + s * 2.0f));
    a.n = last - first;
    int at = 0;
    if (!clamp_mode) {
        at = first % size;
        if (at < 0) at += size;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a.w[k] = keys_weight(fabsf(rcp * (float(first + k) - q)));
        a.i[k] = clamp_mode ? min(max(first + k, 0), size - 1) : at;
        at = at + 1 == size ? 0 : at + 1;
    }
    return a;
}
__device__ __forceinline__ rgba pattern_4x4(const float4 *tex, int width, const axis_taps &cx, const axis_taps &cy)
{
    float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        const float4 *row = tex + cy.i[ky] * width;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            const float wgt = cx.w[kx] * cy.w[ky];
            const float4 c = __ldg(row + cx.i[kx]);
            r = fmaf(wgt, c.x, r); g = fmaf(wgt, c.y, g); b = fmaf(wgt, c.z, b); a = fmaf(wgt, c.w, a);
            wsum += wgt;
        }
    }
    const float norm = 1.0f / wsum;
    return mk(norm * r, norm * g, norm * b, norm * a);
}

// Any other footprint (minification, or a coordinate whose rounded ends give 3 or 5 taps), tap by tap as
// hpp:2296-2328.  Inline on purpose: a call from the compositor's pixel loop would put the warp's 32 pixel
// registers on the stack for the whole kernel.
__device__ __forceinline__ rgba pattern_any(const float4 *tex, int width, int height, bool clamp_mode, float qx, float qy,
                                            float sx, float sy, float rx, float ry)
{
    const int x0 = int(ceilf(qx - sx * 2.0f)), y0 = int(ceilf(qy - sy * 2.0f));
    const int x1 = int(ceilf(qx + sx * 2.0f)), y1 = int(ceilf(qy + sy * 2.0f));
    float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f, wsum = 0.0f;
#pragma unroll 1
    for (int ty = y0; ty < y1; ++ty) {
        const float wy = keys_weight(fabsf(ry * (float(ty) - qy)));
        int yy = ty % height;
        if (yy < 0) yy += height;
        if (clamp_mode) yy = min(max(ty, 0), height - 1);
        const float4 *row = tex + yy * width;
#pragma unroll 1
        for (int tx = x0; tx < x1; ++tx) {
            const float wgt = keys_weight(fabsf(rx * (float(tx) - qx))) * wy;
            int xx = tx % width;
            if (xx < 0) xx += width;
            if (clamp_mode) xx = min(max(tx, 0), width - 1);
            const float4 c = __ldg(row + xx);
            r = fmaf(wgt, c.x, r); g = fmaf(wgt, c.y, g); b = fmaf(wgt, c.z, b); a = fmaf(wgt, c.w, a);
            wsum += wgt;
        }
    }
    const float norm = 1.0f / wsum;
    return mk(norm * r, norm * g, norm * b, norm * a);
}

// Non-solid brushes: kept out of line so the solid path stays small.
__device__ __noinline__ rgba paint_slow(paint_tables f, uint32_t brush, uint32_t draw, float x, float y)
{
    return paint_at(f, f.brushes[brush], f.draws[draw].inverse, v2(x, y));
}

// One WARP owns 8 scanlines x 32 pixels of a tile (8 pixels per lane, in registers) and works completely on its
// own: no block-wide barrier anywhere, so an SM keeps ~20 independent tile-quarters in flight and their dependent
// loads (job table -> tile entry -> sums) overlap.  Per warp:
//   * occlusion culling comes first and costs one load: k_tile_flags left, per tile, the last job that covers the
//     whole tile with an opaque colour (cov = vis = 1, source_over / copy: the draw REPLACES the pixel exactly).
//     Everything before it -- including the old pixels -- is irrelevant; the replay starts at that job;
//   * the old pixels are requested right away otherwise, so that their latency overlaps the job search;
//   * the job search tests 32 candidates per step against the compact table (8 B tile box + flags, 4 B first tile
//     entry) and replays the hits of the step immediately, in order: no hit list, one replay site;
//   * a hit needs three 16-byte loads of its record (warp-uniform address) and ONE coalesced load of this warp's
//     eight rows of the tile entry (lanes 0-7 the sums carried in, 8-15 the first compacted entries, 16-23 the pixel
//     masks), handed to the rows by shuffles -- no shared memory, no warp synchronisation in the solid-colour builds;
//   * per row and pixel: coverage = one popc + one load (tile_cov.cuh), paint, the 4-bit Porter-Duff program.
// Builds of the same code, picked by the host per frame (device_frame::general_compositor):
//   kMode 0  lean: frames that only hold unclipped solid-colour fills and strokes without shadows
//            (the tiger, most UI and plots) -- no gradient/pattern/mask/shadow code
//   kMode 1  + clip masks (in and out) and shadow planes, brushes still solid
//   kMode 2  lean + unclipped gradient brushes with at most kStagedStops stops (full-canvas gradient fills)
//   kMode 3  everything: masks, shadows, gradients, patterns
//   kMode 4  lean + unclipped patterns / images and small gradients (draw_image and pattern fills without clips or shadows)
// kLists: the job search walks the tile row's job list (frames with many jobs) instead of the
// canvas' whole job range (a handful of jobs: the plain loop is leaner).
#ifndef CB200_COMP_CTAS0
#define CB200_COMP_CTAS0 8
#endif
// The 4-bit mix program (hpp:2583-2591) as two affine maps, decoded once per job: mix_fore = f0 + f1 * back.a and
// mix_back = b0 + b1 * fore.a with (f0, f1), (b0, b1) in {(0,0), (0,1), (1,0), (1,-1)}.  fmaf(1, a, 0) = a and
// fmaf(-1, a, 1) = 1 - a exactly, so the per-pixel selects of blend() become two FMAs with the same bits.
struct mix_program { float f0, f1, b0, b1; };
__device__ __forceinline__ mix_program decode_mix(uint32_t op)
{
    mix_program m;
    m.f1 = (op & 1u) ? ((op & 2u) ? -1.0f : 1.0f) : 0.0f;  m.f0 = (op & 2u) ? 1.0f : 0.0f;
    m.b1 = (op & 4u) ? ((op & 8u) ? -1.0f : 1.0f) : 0.0f;  m.b0 = (op & 8u) ? 1.0f : 0.0f;
    return m;
}
// blend_unclipped() with the program decoded (visibility exactly 1)
__device__ __forceinline__ void blend_program(float4 &back, rgba fore, const mix_program &m)
{
    const float mf = fmaf(m.f1, back.w, m.f0), mb = fmaf(m.b1, fore.a, m.b0);
    back = make_float4(fmaf(mf, fore.r, mb * back.x), fmaf(mf, fore.g, mb * back.y), fmaf(mf, fore.b, mb * back.z),
                       fminf(fmaf(mf, fore.a, mb * back.w), 1.0f));
}

// ---- bulk asynchronous copies (the TMA engine's 1-D form: cp.async.bulk, no tensor map) ------------------------
// Used by the kTma build of the compositor: a warp's 8 framebuffer rows (512 contiguous bytes each) travel between
// global and shared memory as eight bulk copies issued by eight lanes, completion through an mbarrier (loads) or a
// bulk group (stores), instead of eight 16-byte LDG/STG per lane.  An A/B against the plain build.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kMode, bool kLists, bool kTma = false>
__global__ void __launch_bounds__(kCompBlock, kMode == 0 ? CB200_COMP_CTAS0 : kMode == 1 ? 7 : kMode == 2 ? 6 : kMode == 3 ? 5 : 4)
k_composite(device_frame f, canvas_target t, int tiles_x, int tile_y0, int eager_load)
{
    constexpr bool kGeneral = kMode == 1 || kMode == 3, kPaint = kMode >= 2, kPattern = kMode >= 3;
    grid_dependency_wait();
    __shared__ __align__(16) staged_brush staged_brushes[kPaint ? kTileWarps : 1];
    frame_header *h = f.hdr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int trow = int(blockIdx.x) / tiles_x;
    const int tx = int(blockIdx.x) - trow * tiles_x, ty = tile_y0 + trow;
    // requests that depend on nothing but the tile go out first: their latencies overlap
    const uint32_t overflow = h->overflow;
    const uint32_t cover = f.tile_cover[blockIdx.x];          // 1 + the last opaque job that covers the whole tile
    const uint32_t search_rows = kLists ? f.row_job_count[trow] : 0u;
    // batches stack their canvases vertically: which canvas is this tile in, and where does it start?
    int canvas = 0, ty_local = ty;
    if (t.n_canvases > 1) {
        const int slot_tiles = t.slot_rows / kTile;
        canvas = ty / slot_tiles;
        ty_local = ty - canvas * slot_tiles;
    }
    const uint2 job_range = t.canvas_jobs[canvas];
    const uint32_t job_end = job_range.x + job_range.y;
    const int row0 = ty_local * kTile + warp * kWarpRows;    // first scanline of this warp, canvas coordinates
    const int x = tx * kTile + lane;
    // scanlines [r_lo, r_hi) of this warp lie in the band; a lane is live when its column is on the canvas
    const int r_lo = max(t.band_y0 - row0, 0), r_hi = min(t.band_y0 + t.band_rows - row0, kWarpRows);
    if (r_hi <= r_lo) return;
    const bool whole = r_lo == 0 && r_hi == kWarpRows && tx * kTile + kTile <= t.width;   // no partial rows or columns
    const uint32_t live_mask = x < t.width ? ((1u << r_hi) - 1u) & ~((1u << r_lo) - 1u) : 0u;
    const uint32_t *row_list = kLists ? f.row_jobs + size_t(trow) * f.row_stride : nullptr;
    const uint32_t search_end = kLists ? search_rows : job_end;
    const size_t local_index = size_t(row0 - t.band_y0) * size_t(t.width) + size_t(x);   // into canvas-sized planes
    float4 *const fb_at = t.fb + (size_t(canvas) * size_t(t.slot_rows) * size_t(t.width) + local_index);
    const size_t pitch = size_t(t.width);
    if (overflow) return;
    if (search_end == (kLists ? 0u : job_range.x) && !t.clear_first) return;      // no job reaches this tile row

    // occlusion culling: the last opaque job that covers this tile voids everything before it -- old pixels included
    const uint32_t first_job = cover ? cover - 1u : job_range.x;
    float4 px[kWarpRows];
    const bool fetch_now = eager_load && !t.clear_first && !cover;
    // kTma: the warp's rows wait in shared memory (bulk copies, one mbarrier per warp) until the first job needs them
    __shared__ __align__(128) float4 tma_rows[kTma ? kTileWarps * kWarpRows * 32 : 1];
    __shared__ __align__(8) unsigned long long tma_bar[kTma ? kTileWarps : 1];
    float4 *const my_rows = tma_rows + (kTma ? warp * kWarpRows * 32 : 0);
    const uint32_t my_bar = smem_u32(&tma_bar[kTma ? warp : 0]);
    bool pending = false;
    if (kTma) {
        if (lane == 0) mbar_init(my_bar, 1);
        __syncwarp();
    }
    auto settle = [&]() {                                      // the rows a bulk load brought: shared memory -> registers
        if (kTma && pending) {
            mbar_wait(my_bar, 0);
#pragma unroll
            for (int r = 0; r < kWarpRows; ++r) px[r] = my_rows[r * 32 + lane];
            pending = false;
        }
    };
    auto fetch = [&]() {
        if (kTma && whole) {
            if (lane == 0) mbar_expect_tx(my_bar, kWarpRows * 512);
            __syncwarp();
            if (lane < kWarpRows) bulk_load(smem_u32(my_rows + lane * 32), fb_at - lane + size_t(lane) * pitch, 512, my_bar);
            pending = true;
        } else if (whole) {
#pragma unroll
            for (int r = 0; r < kWarpRows; ++r) px[r] = __ldcs(fb_at + size_t(r) * pitch);
        } else {
#pragma unroll
            for (int r = 0; r < kWarpRows; ++r)
                if (live_mask >> r & 1u) px[r] = __ldcs(fb_at + size_t(r) * pitch);
        }
    };
    if (!(fetch_now && whole) || kTma) {
#pragma unroll
        for (int r = 0; r < kWarpRows; ++r) px[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (fetch_now) fetch();
    bool loaded = fetch_now || t.clear_first != 0 || cover != 0;
    bool touched = cover != 0;
    uint32_t painted = 0;
    const paint_tables tables = { f.colors, f.stops, f.texels, f.brushes, f.draws };
    staged_brush &sbrush = staged_brushes[kPaint ? warp : 0];

    for (uint32_t base = kLists ? 0u : first_job; base < search_end; base += 32) {
        const uint32_t at = base + uint32_t(lane);
        const uint32_t j = kLists ? (at < search_end ? row_list[at] : job_end) : at;
        uint32_t te = 0;
        bool hit = false;
        if (j < job_end && j >= first_job) {
            const uint2 box = f.job_box[j];
            const uint32_t te_base = f.job_te[j];
            const int bx0 = int(box.x & 0x7ffu), by0 = int((box.x >> 11) & 0x7ffu);
            const int bx1 = int((box.x >> 22) & 0x3ffu) | int((box.y & 1u) << 10), by1 = int((box.y >> 1) & 0x7ffu);
            if (ty_local >= by0 && ty_local <= by1 && tx >= bx0) {
                if (tx <= bx1) {
                    if ((box.y >> 12 & 3u) == JOB_SHADOW) hit = true;
                    else {
                        te = te_base + uint32_t(ty_local - by0) * uint32_t(bx1 - bx0 + 1) + uint32_t(tx - bx0);
                        hit = (box.y & JOBBOX_EVERYWHERE) || (f.te_flags[te] & TE_NONEMPTY);
                    }
                } else if ((box.y & JOBBOX_LEAKY) && bx1 >= bx0) {
                    // to the right of a job with scanlines whose coverage never returns to zero (leak_rec): no tile
                    // entry there, the rows come from the frame's leak list
                    hit = true;
                    te = kNoRun;
                }
            }
        }
        if (hit) {
            // the hits of this step are replayed one after the other; ask for what each replay starts with -- the job's
            // record and this warp's rows of its tile entry -- now, so that those loads find their lines in L1
            prefetch_l1(f.comp + j);
            if (te != kNoRun && (f.job_box[j].y >> 12 & 3u) != JOB_SHADOW) {
                const uint32_t slot = te * kTile + uint32_t(warp * kWarpRows);
                prefetch_l1(f.te_backdrop + slot); prefetch_l1(f.te_first + slot); prefetch_l1(f.te_mask + slot);
            }
        }
        uint32_t votes = __ballot_sync(0xffffffffu, hit);
        if (votes && !loaded) {                               // lazy variant: fetch the old pixels on first use
            fetch();
            loaded = true;
        }
        touched = touched || votes != 0;
        if (votes) settle();
        while (votes) {
            const int src = __ffs(int(votes)) - 1;
            votes &= votes - 1u;
            const uint32_t jj = __shfl_sync(0xffffffffu, j, src), tte = __shfl_sync(0xffffffffu, te, src);
            // the job's record: three 16-byte loads, the same address in every lane
            const comp_rec *rec = f.comp + jj;
            const uint4 head = __ldg(reinterpret_cast<const uint4 *>(rec));            // kind, op, flags, brush_type
            const float4 colour = __ldg(reinterpret_cast<const float4 *>(rec) + 1);
            const uint4 tail = __ldg(reinterpret_cast<const uint4 *>(rec) + 2);        // alpha, mask_src, mask_dst, brush
            const uint32_t kind = head.x, op = head.y, brush_type = head.w;
            const float alpha = __uint_as_float(tail.x);
            // this warp's eight rows of the tile entry in one load: lanes 0-7 carried-in sums, 8-15 first compacted
            // entries, 16-23 pixel masks
            const bool leak_tile = tte == kNoRun;
            uint32_t info = (lane >= 8 && lane < 16) ? kNoRun : 0u;
            if (kind != JOB_SHADOW && !leak_tile && lane < 24) {
                const uint32_t slot = tte * kTile + uint32_t(warp * kWarpRows + (lane & 7));
                info = lane < 8 ? __float_as_uint(f.te_backdrop[slot]) : lane < 16 ? f.te_first[slot] : f.te_mask[slot];
            }
            if (leak_tile) {
                // rare: the rows of this job that leak are few, the list is short -- every lane scans a share and
                // hands what it finds to the lane that holds that row's carried-in sum
                for (uint32_t k0 = 0; k0 < h->n_leaks; k0 += 32) {
                    const uint32_t k = k0 + uint32_t(lane);
                    leak_rec l = { 0xffffffffu, 0, 0.0f, 0u };
                    if (k < h->n_leaks) l = f.leaks[k];
                    const bool mine = l.job == jj && l.y >= row0 && l.y < row0 + kWarpRows;
                    uint32_t found = __ballot_sync(0xffffffffu, mine);
                    while (found) {
                        const int from = __ffs(int(found)) - 1;
                        found &= found - 1u;
                        const int r = __shfl_sync(0xffffffffu, l.y, from) - row0;
                        const float sum = __shfl_sync(0xffffffffu, l.sum, from);
                        if (lane == r) info = __float_as_uint(sum);
                    }
                }
            }
            // rows with an edge inside the tile (most tiles of a filled shape have none: the carried-in sum is all)
            const uint32_t edge_rows = (__ballot_sync(0xffffffffu, info != kNoRun) >> 8) & 0xffu;
            const float *mask = (kGeneral && tail.y) ? t.mask_planes[tail.y] : nullptr;
            if (kGeneral && kind == JOB_SHADOW) {
                const int4 box = __ldg(reinterpret_cast<const int4 *>(rec) + 3);       // cx0, cy0, cx1, cy1
                const int4 place = __ldg(reinterpret_cast<const int4 *>(rec) + 4);     // border, left, top, pitch
                const uint2 off = __ldg(reinterpret_cast<const uint2 *>(rec) + 10);    // plane offset
                const float *plane = f.planes + (uint64_t(off.y) << 32 | off.x);
                const rgba tint = mk(colour.x, colour.y, colour.z, colour.w);
                // the eight rows' plane samples are requested together (one dependent load per row, consumed right
                // away, left 59 % of this build's stall samples on that load: config 3)
                const bool x_ok = x >= box.x && x < box.z;
                const float *column = plane + ptrdiff_t(row0 + place.x - place.z) * ptrdiff_t(place.w) + ptrdiff_t(x + place.x - place.y);
                float s[kWarpRows];
                uint32_t rows_in = 0;
#pragma unroll
                for (int r = 0; r < kWarpRows; ++r) {
                    const int y = row0 + r;
                    const bool in = (live_mask >> r & 1u) && x_ok && y >= box.y && y < box.w;
                    s[r] = in ? __ldg(column + r * place.w) : 0.0f;
                    rows_in |= uint32_t(in) << r;
                }
                if (!mask) {
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        if (!(rows_in >> r & 1u)) continue;
                        blend_unclipped(px[r], scale(alpha * s[r], tint), op);
                        ++painted;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        if (!(rows_in >> r & 1u)) continue;
                        const float vis = fminf(fabsf(mask[size_t(row0 + r - t.band_y0) * pitch + size_t(x)]), 1.0f);
                        if (vis < kThreshold) continue;
                        blend(px[r], scale(alpha * s[r], tint), op, vis);
                        ++painted;
                    }
                }
                continue;
            }
            const bool everywhere = (~op & 8u) != 0;
            const rgba flat = mk(colour.x, colour.y, colour.z, colour.w);
            float *mask_out = (kGeneral && kind == JOB_CLIP) ? t.mask_planes[tail.z] : nullptr;
            const bool gradient = brush_type == CB200_BRUSH_LINEAR || brush_type == CB200_BRUSH_RADIAL;
            bool staged = false;
            if (kPaint && (gradient || (kPattern && brush_type == CB200_BRUSH_PATTERN)) && kind == JOB_MAIN) {
                // stage the brush: record (14 words), brush-space matrix (6 words), gradient stops
                const brush_rec *gb = &f.brushes[tail.w];
                const uint32_t n_stops = gradient ? gb->n_colors : 0;
                staged = n_stops <= kStagedStops;
                __syncwarp();
                if (staged) {
                    if (lane < int(sizeof(brush_rec) / 4))
                        reinterpret_cast<uint32_t *>(&sbrush.b)[lane] = reinterpret_cast<const uint32_t *>(gb)[lane];
                    if (lane < 6)
                        reinterpret_cast<float *>(&sbrush.inv)[lane] = reinterpret_cast<const float *>(&f.draws[rec->draw].inverse)[lane];
                    if (uint32_t(lane) < n_stops) {
                        const float stop = f.stops[gb->first_color + lane];
                        sbrush.stops[lane] = stop;
                        sbrush.inv_span[lane] = lane ? 1.0f / (stop - f.stops[gb->first_color + lane - 1]) : 0.0f;
                        sbrush.colors[lane] = f.colors[gb->first_color + lane];
                    }
                }
                __syncwarp();
            }
            // coverage sum of this lane's pixel in row r (tile_cov.cuh); `r` is a compile-time constant at every use
            auto row_sum = [&](int r) -> float {
                const float carried = __uint_as_float(__shfl_sync(0xffffffffu, info, r));
                if (!(edge_rows >> r & 1u)) return carried;
                return pixel_sum<false>(f.cumulative, carried, __shfl_sync(0xffffffffu, info, 8 + r), __shfl_sync(0xffffffffu, info, 16 + r));
            };
            if ((!kGeneral && !kPaint) || (!mask && !mask_out && brush_type == CB200_BRUSH_COLOR)) {
                // the common case -- unclipped solid colour.  Dead lanes and rows are blended too (they hold zeros and
                // are never stored): no per-row liveness test here.
                const float scale_all = alpha;
                if (op == 14u) {
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                        if (!(cov >= kThreshold)) continue;
                        ++painted;
                        blend_unclipped(px[r], scale(cov * scale_all, flat), 14u);
                    }
                } else {
                    const mix_program m = decode_mix(op);
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                        if (!(cov >= kThreshold || everywhere)) continue;
                        ++painted;
                        blend_program(px[r], scale(cov * scale_all, flat), m);
                    }
                }
            } else if (kPaint && staged && gradient && !mask && !mask_out) {
                // unclipped gradient fill: the brush set-up is hoisted out of the pixel loop
                const gradient_ctx g = make_gradient_ctx(sbrush);
                const mix_program m = decode_mix(op);
#pragma unroll
                for (int r = 0; r < kWarpRows; ++r) {
                    const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                    if (!(cov >= kThreshold || everywhere)) continue;
                    ++painted;
                    rgba paint = gradient_at(g, sbrush, float(x) + 0.5f, float(row0 + r) + 0.5f);
                    blend_program(px[r], scale(cov * alpha, paint), m);
                }
            } else if (kPattern && staged && brush_type == CB200_BRUSH_PATTERN && !mask && !mask_out) {
                // unclipped pattern / image fill (hpp:2274-2330)
                const affine inv = sbrush.inv;
                const int tex_w = sbrush.b.width, tex_h = sbrush.b.height;
                const uint32_t repetition = sbrush.b.repetition;
                const bool clamp_mode = (sbrush.b.flags & CB200_BRUSH_CLAMP) != 0;
                const float4 *tex = f.texels + sbrush.b.texel_offset;
                const float w = float(tex_w), hgt = float(tex_h);
                const float sx = fmaxf(1.0f, fminf(fabsf(inv.a) + fabsf(inv.c), w * 0.25f));
                const float sy = fmaxf(1.0f, fminf(fabsf(inv.b) + fabsf(inv.d), hgt * 0.25f));
                const float rx = 1.0f / sx, ry = 1.0f / sy;
                const mix_program m = decode_mix(op);
                // Rows of this lane whose pixel goes tap by tap (pattern_any): all of them under a rotated or skewed
                // brush matrix, otherwise only pixels whose footprint is not four taps wide.
                uint32_t slow_rows = 0;
                if (inv.b == 0.0f && inv.c == 0.0f) {
                    // Axis-aligned matrix (every draw_image without rotation): the footprint is a product.  A lane's
                    // column taps hold for all of its rows; lane r < 8 prepares the row taps of scanline r and leaves
                    // them in shared memory; the four texel rows are filtered along x once per lane (row_filter) and
                    // shared by all pixel rows that need them -- at 8x magnification a warp's 8 scanlines touch 4 or 5
                    // texel rows, i.e. 16-20 texel loads per lane instead of 128.  Sum_y wy (Sum_x wx c) /
                    // (Sum wx Sum wy): the reference's sum in another order (a few ulp; inside the promised 1e-4).
                    const vec2 p_col = apply(inv, v2(float(x) + 0.5f, float(row0) + 0.5f));
                    const axis_taps cx = keys_taps(p_col.x - 0.5f, sx, rx, tex_w, clamp_mode);
                    const bool x_bad = cx.n != 4;
                    const bool x_out = (repetition & 2u) && (p_col.x < 0.0f || w <= p_col.x);
                    const float sum_wx = ((cx.w[0] + cx.w[1]) + cx.w[2]) + cx.w[3];
                    __syncwarp();
                    if (lane < kWarpRows) {
                        const vec2 p_row = apply(inv, v2(float(x) + 0.5f, float(row0 + lane) + 0.5f));
                        const axis_taps cy = keys_taps(p_row.y - 0.5f, sy, ry, tex_h, clamp_mode);
                        const bool y_out = (repetition & 1u) && (p_row.y < 0.0f || hgt <= p_row.y);
                        sbrush.row_w[lane] = make_float4(cy.w[0], cy.w[1], cy.w[2], cy.w[3]);
                        sbrush.row_i[lane] = make_int4(cy.i[0] * tex_w, cy.i[1] * tex_w, cy.i[2] * tex_w, cy.i[3] * tex_w);
                        sbrush.row_flags[lane] = (cy.n != 4 ? 1 : 0) | (y_out ? 2 : 0);
                    }
                    __syncwarp();
                    rgba filtered[4];
                    int4 filtered_at = make_int4(-1, -1, -1, -1);
                    auto row_filter = [&](int texel_row_offset) -> rgba {
                        const float4 *row = tex + texel_row_offset;
                        float fr = 0.0f, fg = 0.0f, fb = 0.0f, fa = 0.0f;
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx) {
                            const float4 c = __ldg(row + cx.i[kx]);
                            fr = fmaf(cx.w[kx], c.x, fr); fg = fmaf(cx.w[kx], c.y, fg);
                            fb = fmaf(cx.w[kx], c.z, fb); fa = fmaf(cx.w[kx], c.w, fa);
                        }
                        return mk(fr, fg, fb, fa);
                    };
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                        if (!(cov >= kThreshold || everywhere)) continue;
                        ++painted;
                        const int flags = sbrush.row_flags[r];
                        if (x_out || (flags & 2)) {
                            blend_program(px[r], mk(0.0f, 0.0f, 0.0f, 0.0f), m);
                            continue;
                        }
                        if (x_bad || (flags & 1)) {
                            slow_rows |= 1u << r;
                            continue;
                        }
                        const float4 wy = sbrush.row_w[r];
                        const int4 at = sbrush.row_i[r];
                        if (!(at.x == filtered_at.x && at.y == filtered_at.y && at.z == filtered_at.z && at.w == filtered_at.w)) {
                            if (at.x == filtered_at.y && at.y == filtered_at.z && at.z == filtered_at.w) {
                                filtered[0] = filtered[1]; filtered[1] = filtered[2]; filtered[2] = filtered[3];
                            } else {
                                filtered[0] = row_filter(at.x); filtered[1] = row_filter(at.y); filtered[2] = row_filter(at.z);
                            }
                            filtered[3] = row_filter(at.w);
                            filtered_at = at;
                        }
                        float pr = wy.x * filtered[0].r, pg = wy.x * filtered[0].g, pb = wy.x * filtered[0].b, pa = wy.x * filtered[0].a;
                        pr = fmaf(wy.y, filtered[1].r, pr); pg = fmaf(wy.y, filtered[1].g, pg); pb = fmaf(wy.y, filtered[1].b, pb); pa = fmaf(wy.y, filtered[1].a, pa);
                        pr = fmaf(wy.z, filtered[2].r, pr); pg = fmaf(wy.z, filtered[2].g, pg); pb = fmaf(wy.z, filtered[2].b, pb); pa = fmaf(wy.z, filtered[2].a, pa);
                        pr = fmaf(wy.w, filtered[3].r, pr); pg = fmaf(wy.w, filtered[3].g, pg); pb = fmaf(wy.w, filtered[3].b, pb); pa = fmaf(wy.w, filtered[3].a, pa);
                        // cov * alpha / total weight in one factor
                        const float norm = (cov * alpha) * __frcp_rn(sum_wx * (((wy.x + wy.y) + wy.z) + wy.w));
                        blend_program(px[r], mk(norm * pr, norm * pg, norm * pb, norm * pa), m);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kWarpRows; ++r) {
                        const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                        if (cov >= kThreshold || everywhere) { slow_rows |= 1u << r; ++painted; }
                    }
                }
                if (__any_sync(0xffffffffu, slow_rows != 0)) {
                    // one copy of the tap-by-tap code: the pixel registers rotate through px[0] instead of the loop
                    // being unrolled eight times
#pragma unroll 1
                    for (int r = 0; r < kWarpRows; ++r) {
                        float4 cur = px[0];
                        const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                        if (slow_rows >> r & 1u) {
                            const vec2 p = apply(inv, v2(float(x) + 0.5f, float(row0 + r) + 0.5f));
                            rgba paint = mk(0.0f, 0.0f, 0.0f, 0.0f);
                            if (!(((repetition & 2u) && (p.x < 0.0f || w <= p.x)) || ((repetition & 1u) && (p.y < 0.0f || hgt <= p.y))))
                                paint = pattern_any(tex, tex_w, tex_h, clamp_mode, p.x - 0.5f, p.y - 0.5f, sx, sy, rx, ry);
                            blend_program(cur, scale(cov * alpha, paint), m);
                        }
#pragma unroll
                        for (int k = 0; k + 1 < kWarpRows; ++k) px[k] = px[k + 1];
                        px[kWarpRows - 1] = cur;
                    }
                }
            } else if (kGeneral) {
#pragma unroll
                for (int r = 0; r < kWarpRows; ++r) {
                    const int y = row0 + r;
                    const float cov = fminf(fabsf(row_sum(r)), 1.0f);
                    if (!(live_mask >> r & 1u)) continue;
                    const size_t at_px = local_index + size_t(r) * pitch;
                    float vis = mask ? fminf(fabsf(mask[at_px]), 1.0f) : 1.0f;
                    if (mask_out) { mask_out[at_px] = cov * vis; continue; }
                    if (!((cov >= kThreshold || everywhere) && vis >= kThreshold)) continue;
                    ++painted;
                    rgba paint;
                    if (brush_type == CB200_BRUSH_COLOR) paint = flat;
                    else if (brush_type == 0xffu) paint = mk(0.0f, 0.0f, 0.0f, 0.0f);
                    else if (!kPattern) paint = mk(0.0f, 0.0f, 0.0f, 0.0f);     // unreachable: the host picked mode 3
                    else if (staged && gradient) paint = paint_gradient(sbrush, float(x) + 0.5f, float(y) + 0.5f);
                    else if (staged) paint = paint_pattern(f.texels, &sbrush, float(x) + 0.5f, float(y) + 0.5f);
                    else paint = paint_slow(tables, tail.w, rec->draw, float(x) + 0.5f, float(y) + 0.5f);
                    blend(px[r], scale(cov * alpha, paint), op, vis);
                }
            }
        }
    }
    settle();                                                // (a bulk load must land before its shared memory goes away)
    if (!touched && !t.clear_first) return;                  // no job reaches these pixels: leave them alone
    if (kTma && whole) {
#pragma unroll
        for (int r = 0; r < kWarpRows; ++r) my_rows[r * 32 + lane] = px[r];
        fence_async_proxy();
        __syncwarp();
        if (lane < kWarpRows) {
            bulk_store(fb_at - lane + size_t(lane) * pitch, smem_u32(my_rows + lane * 32), 512);
            bulk_store_commit_and_wait_read();
        }
    } else if (whole) {
#pragma unroll
        for (int r = 0; r < kWarpRows; ++r) fb_at[size_t(r) * pitch] = px[r];
    } else {
#pragma unroll
        for (int r = 0; r < kWarpRows; ++r)
            if (live_mask >> r & 1u) fb_at[size_t(r) * pitch] = px[r];
    }
    // statistics: composited pixel count of the frame (dead lanes of edge tiles included in the solid-colour path)
    painted = __reduce_add_sync(0xffffffffu, painted);
    if (lane == 0 && painted) atomicAdd(&h->composited_slots[blockIdx.x & 31u], (unsigned long long)painted);
}

// One warp per tile row of the target: the ordered list of jobs whose composite box reaches the row.
__global__ void __launch_bounds__(kBlock) k_row_lists(device_frame f, canvas_target t, int tile_y0, int n_rows)
{
    grid_dependency_wait();
    if (f.hdr->overflow) return;
    const int lane = threadIdx.x & 31;
    const int row = int((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int ty = tile_y0 + row;
    int canvas = 0, ty_local = ty;
    if (t.n_canvases > 1) {
        const int slot_tiles = t.slot_rows / kTile;
        canvas = ty / slot_tiles;
        ty_local = ty - canvas * slot_tiles;
    }
    const uint2 range = t.canvas_jobs[canvas];
    uint32_t *list = f.row_jobs + size_t(row) * f.row_stride;
    uint32_t n = 0;
    for (uint32_t base = range.x; base < range.x + range.y; base += 32) {
        const uint32_t j = base + uint32_t(lane);
        bool in = false;
        if (j < range.x + range.y) {
            const uint2 box = f.job_box[j];
            const int by0 = int((box.x >> 11) & 0x7ffu), by1 = int((box.y >> 1) & 0x7ffu);
            const int bx0 = int(box.x & 0x7ffu), bx1 = int((box.x >> 22) & 0x3ffu) | int((box.y & 1u) << 10);
            in = bx1 >= bx0 && ty_local >= by0 && ty_local <= by1;
        }
        const uint32_t votes = __ballot_sync(0xffffffffu, in);
        if (in) list[n + __popc(votes & ((1u << lane) - 1u))] = j;
        n += __popc(votes);
    }
    if (lane == 0) f.row_job_count[row] = n;
}

}  // namespace

void launch_composite(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    int tiles_x = (t.width + kTile - 1) / kTile;
    int ty0 = t.band_y0 / kTile, ty1 = (t.band_y0 + t.band_rows - 1) / kTile;
    if (t.n_canvases > 1) { ty0 = 0; ty1 = t.n_canvases * (t.slot_rows / kTile) - 1; }
    int tiles = tiles_x * (ty1 - ty0 + 1);
    // Eager: request the old pixels before the job search (hides their latency).  Tiles that an opaque job covers
    // are known up front (tile_cover) and never read.  CB200_EAGER_LOAD=0 loads on first use instead.
    int eager = 1;
    if (const char *e = getenv("CB200_EAGER_LOAD")) eager = atoi(e);
    // CB200_TMA=1: lean build whose framebuffer rows move as bulk asynchronous copies (measured A/B)
    static const bool tma = [] { const char *e = getenv("CB200_TMA"); return e && atoi(e) != 0; }();
    if (f.row_jobs) {
        const int n_rows = ty1 - ty0 + 1;
        launch_pdl(k_row_lists, (n_rows * 32 + kBlock - 1) / kBlock, kBlock, 0, s, f, t, ty0, n_rows);
    }
    auto go = [&](auto kernel) { launch_pdl(kernel, tiles, kCompBlock, 0, s, f, t, tiles_x, ty0, eager); };
    if (f.row_jobs) {
        if (f.general_compositor == 4) go(k_composite<4, true>);
        else if (f.general_compositor == 3) go(k_composite<3, true>);
        else if (f.general_compositor == 2) go(k_composite<2, true>);
        else if (f.general_compositor == 1) go(k_composite<1, true>);
        else if (tma) go(k_composite<0, true, true>);
        else go(k_composite<0, true>);
    } else {
        if (f.general_compositor == 4) go(k_composite<4, false>);
        else if (f.general_compositor == 3) go(k_composite<3, false>);
        else if (f.general_compositor == 2) go(k_composite<2, false>);
        else if (f.general_compositor == 1) go(k_composite<1, false>);
        else if (tma) go(k_composite<0, false, true>);
        else go(k_composite<0, false>);
    }
}

}  // namespace cb200
