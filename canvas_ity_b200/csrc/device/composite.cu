// composite.cu -- K7: tile compositor.  One CTA owns one 32x32 framebuffer tile,
// keeps its 1024 linear premultiplied float RGBA pixels in registers (4 per
// thread), replays every job that touches the tile IN SUBMISSION ORDER and stores
// the tile once.  Per job and pixel it does what the reference does per span
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

// paint_pixel, hpp:2265-2377.  `at` is the device-space pixel centre.
__device__ rgba paint_at(const device_frame &f, const brush_rec &b, const affine &inv, vec2 at)
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

// the mix program of hpp:2583-2591
__device__ __forceinline__ void blend(float4 &back, rgba fore, uint32_t op, float vis)
{
    float mf = (op & 1u) ? back.w : 0.0f;
    if (op & 2u) mf = 1.0f - mf;
    float mb = (op & 4u) ? fore.a : 0.0f;
    if (op & 8u) mb = 1.0f - mb;
    float r = mf * fore.r + mb * back.x, g = mf * fore.g + mb * back.y, b = mf * fore.b + mb * back.z;
    float a = fminf(mf * fore.a + mb * back.w, 1.0f);
    float keep = 1.0f - vis;
    back = make_float4(vis * r + keep * back.x, vis * g + keep * back.y, vis * b + keep * back.z,
                       vis * a + keep * back.w);
}

constexpr int kRowsPerThread = kTile * kTile / kBlock;      // 4

__global__ void __launch_bounds__(kBlock) k_composite(device_frame f, canvas_target t, int sb,
                                                       int tiles_x, int tile_y0)
{
    __shared__ uint32_t sm[33];
    __shared__ uint32_t job_list[kBlock];
    __shared__ float row_buf[kBlock / 32][kTile];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x % tiles_x, ty = tile_y0 + blockIdx.x / tiles_x;
    const int x = tx * kTile + lane;
    const int band_y1 = t.band_y0 + t.band_rows;
    const bool x_in = x < t.width;

    float4 px[kRowsPerThread];
    int py[kRowsPerThread];
    bool live[kRowsPerThread];
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
        py[k] = ty * kTile + warp + k * (kBlock / 32);
        live[k] = x_in && py[k] >= t.band_y0 && py[k] < band_y1;
        px[k] = live[k] ? t.fb[size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x)]
                        : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const cov_source cs = make_cov_source(f, sb);
    const int tile_x0 = tx * kTile, tile_y0p = ty * kTile;
    const uint32_t n_jobs = h->n_jobs;
    unsigned long long painted = 0;

    for (uint32_t base = 0; base < n_jobs; base += kBlock) {
        // which of these 256 jobs touch this tile?  (ordered compaction)
        uint32_t j = base + threadIdx.x, hit = 0;
        if (j < n_jobs) {
            const job_rec &jr = f.jobs[j];
            if (jr.cx0 < tile_x0 + kTile && jr.cx1 > tile_x0 && jr.cy0 < tile_y0p + kTile &&
                jr.cy1 > tile_y0p) {
                if (jr.kind == JOB_SHADOW) hit = 1;
                else {
                    uint32_t te = jr.te_base + uint32_t(ty - jr.ty0) * uint32_t(jr.tw) + uint32_t(tx - jr.tx0);
                    bool everywhere = jr.kind == JOB_CLIP || (~f.draws[jr.draw].op & 8u);
                    hit = (everywhere || f.te_flags[te]) ? 1 : 0;
                }
            }
        }
        uint32_t n_hit;
        uint32_t slot = block_exclusive_scan(hit, sm, n_hit);
        if (hit) job_list[slot] = j;
        __syncthreads();

        for (uint32_t q = 0; q < n_hit; ++q) {
            const uint32_t jj = job_list[q];
            const job_rec &jr = f.jobs[jj];
            const draw_rec &d = f.draws[jr.draw];
            const float *mask = d.mask_src ? t.mask_planes[d.mask_src] : nullptr;
            if (jr.kind == JOB_SHADOW) {
                const float *plane = f.planes + jr.plane_offset;
                const rgba tint = mk(d.shadow_color[0], d.shadow_color[1], d.shadow_color[2], d.shadow_color[3]);
#pragma unroll
                for (int k = 0; k < kRowsPerThread; ++k) {
                    if (!live[k] || x < jr.cx0 || x >= jr.cx1 || py[k] < jr.cy0 || py[k] >= jr.cy1) continue;
                    float vis = mask ? fminf(fabsf(mask[size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x)]), 1.0f) : 1.0f;
                    if (vis < kThreshold) continue;
                    float s = plane[size_t(py[k] + jr.border - jr.top) * size_t(jr.bw) + size_t(x + jr.border - jr.left)];
                    blend(px[k], scale(d.global_alpha * s, tint), d.op, vis);
                    ++painted;
                }
                continue;
            }
            const uint32_t te = jr.te_base + uint32_t(ty - jr.ty0) * uint32_t(jr.tw) + uint32_t(tx - jr.tx0);
            const brush_rec *br = jr.kind == JOB_CLIP ? nullptr : &f.brushes[d.brush];
            const bool everywhere = (~d.op & 8u) != 0;
            float *mask_out = jr.kind == JOB_CLIP ? t.mask_planes[d.mask_dst] : nullptr;
#pragma unroll
            for (int k = 0; k < kRowsPerThread; ++k) {
                const int ly = warp + k * (kBlock / 32);
                // warp-uniform: every lane of the warp shares the row
                float sum = tile_row_sum(cs, te, ly, jj, py[k], tile_x0, row_buf[warp]);
                float cov = fminf(fabsf(sum), 1.0f);
                if (!live[k]) continue;
                size_t at = size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x);
                float vis = mask ? fminf(fabsf(mask[at]), 1.0f) : 1.0f;
                if (mask_out) { mask_out[at] = cov * vis; continue; }
                if (!((cov >= kThreshold || everywhere) && vis >= kThreshold)) continue;
                rgba paint = paint_at(f, *br, d.inverse, v2(float(x) + 0.5f, float(py[k]) + 0.5f));
                blend(px[k], scale(cov * d.global_alpha, paint), d.op, vis);
                ++painted;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k)
        if (live[k]) t.fb[size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x)] = px[k];
    // statistics: composited pixel count of the frame
    for (int off = 16; off; off >>= 1) painted += __shfl_down_sync(0xffffffffu, painted, off);
    if (lane == 0 && painted) atomicAdd(&h->composited_pixels, painted);
}

}  // namespace

void launch_composite(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    int tiles_x = (t.width + kTile - 1) / kTile;
    int ty0 = t.band_y0 / kTile, ty1 = (t.band_y0 + t.band_rows - 1) / kTile;
    int tiles = tiles_x * (ty1 - ty0 + 1);
    k_composite<<<tiles, kBlock, 0, s>>>(f, t, sorted_buffer, tiles_x, ty0);
}

}  // namespace cb200
