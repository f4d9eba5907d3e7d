// shadow.cu -- K9: drop shadows (reference render_shadow, hpp:2395-2539).
//
//   k_shadow_raster  per 32x32 tile of a shadow job's working rectangle (padded
//                    canvas space): coverage * paint.alpha -> float plane
//                    (hpp:2430-2452), coverage rebuilt from the sorted runs exactly
//                    like the compositor does (tile_cov.cuh)
//   k_blur_pass      one extended-box pass (Gwosdek et al.) along rows or columns,
//                    zero outside the working rectangle; three passes per axis
//                    (hpp:2453-2503).  Each output is the direct windowed sum
//                      (w1+w2) * sum_{|d|<=r} s[i+d] + w1 * (s[i-r-1] + s[i+r+1])
//                    which is what the reference's running sum maintains.
// The blurred plane is consumed by the tile compositor (composite.cu).
//
// Round-1 shape: six streaming passes that ping-pong between two planes through
// L2; fusing the three passes of an axis in shared memory is the next step for
// this kernel.
#include "frame.cuh"
#include "tile_cov.cuh"

namespace cb200 {

namespace {

__device__ __forceinline__ float keys_weight(float t)
{
    return t < 1.0f ? (1.5f * t - 2.5f) * t * t + 1.0f : ((-0.5f * t + 2.5f) * t - 4.0f) * t + 2.0f;
}

// alpha of paint_pixel (hpp:2265-2377); shadows only use .a (hpp:2443-2445)
__device__ float paint_alpha(const device_frame &f, const brush_rec &b, const affine &inv, vec2 at)
{
    if (b.n_colors == 0) return 0.0f;
    if (b.type == CB200_BRUSH_COLOR) return f.colors[b.first_color].w;
    vec2 p = apply(inv, at);
    if (b.type == CB200_BRUSH_PATTERN) {
        float w = float(b.width), h = float(b.height);
        if (((b.repetition & 2u) && (p.x < 0.0f || w <= p.x)) ||
            ((b.repetition & 1u) && (p.y < 0.0f || h <= p.y)))
            return 0.0f;
        float sx = fabsf(inv.a) + fabsf(inv.c), sy = fabsf(inv.b) + fabsf(inv.d);
        sx = fmaxf(1.0f, fminf(sx, w * 0.25f));
        sy = fmaxf(1.0f, fminf(sy, h * 0.25f));
        float rx = 1.0f / sx, ry = 1.0f / sy;
        p = p - v2(0.5f, 0.5f);
        int x0 = int(ceilf(p.x - sx * 2.0f)), y0 = int(ceilf(p.y - sy * 2.0f));
        int x1 = int(ceilf(p.x + sx * 2.0f)), y1 = int(ceilf(p.y + sy * 2.0f));
        const float4 *tex = f.texels + b.texel_offset;
        const bool clamp_mode = (b.flags & CB200_BRUSH_CLAMP) != 0;
        float acc = 0.0f, wsum = 0.0f;
        for (int ty = y0; ty < y1; ++ty) {
            float wy = keys_weight(fabsf(ry * (float(ty) - p.y)));
            int yy = ty % b.height;
            if (yy < 0) yy += b.height;
            if (clamp_mode) yy = min(max(ty, 0), b.height - 1);
            for (int tx = x0; tx < x1; ++tx) {
                float wx = keys_weight(fabsf(rx * (float(tx) - p.x)));
                int xx = tx % b.width;
                if (xx < 0) xx += b.width;
                if (clamp_mode) xx = min(max(tx, 0), b.width - 1);
                float wgt = wx * wy;
                acc += wgt * tex[size_t(yy) * size_t(b.width) + size_t(xx)].w;
                wsum += wgt;
            }
        }
        return (1.0f / wsum) * acc;
    }
    vec2 rel = p - v2(b.sx, b.sy), axis = v2(b.ex, b.ey) - v2(b.sx, b.sy);
    float along = dot(rel, axis), axis2 = dot(axis, axis);
    float t;
    if (b.type == CB200_BRUSH_LINEAR) {
        if (axis2 == 0.0f) return 0.0f;
        t = along / axis2;
    } else {
        float dr = b.r1 - b.r0;
        float qa = axis2 - dr * dr;
        float qb = -2.0f * (along + b.r0 * dr);
        float qc = dot(rel, rel) - b.r0 * b.r0;
        float disc = qb * qb - 4.0f * qa * qc;
        if (disc < 0.0f || (axis2 == 0.0f && dr == 0.0f)) return 0.0f;
        float root = sqrtf(disc), inv2a = 1.0f / (2.0f * qa);
        float ta = (-qb - root) * inv2a, tb = (-qb + root) * inv2a;
        if (b.r0 + dr * tb >= 0.0f) t = tb;
        else if (b.r0 + dr * ta >= 0.0f) t = ta;
        else return 0.0f;
    }
    const float *stops = f.stops + b.first_color;
    uint32_t hi = 0;
    while (hi < b.n_colors && !(t < stops[hi])) ++hi;
    if (hi == 0) return f.colors[b.first_color].w;
    if (hi == b.n_colors) return f.colors[b.first_color + b.n_colors - 1].w;
    float m = (t - stops[hi - 1]) / (stops[hi] - stops[hi - 1]);
    float lo = f.colors[b.first_color + hi - 1].w, up = f.colors[b.first_color + hi].w;
    return lo + m * (up - lo);
}

// grid: (tile stride, shadow job)
__global__ void __launch_bounds__(kBlock) k_shadow_raster(device_frame f, int sb)
{
    __shared__ float row_buf[kBlock / 32][kTile];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t j = f.shadow_jobs[blockIdx.y];
    const job_rec &jr = f.jobs[j];
    const uint32_t tiles = uint32_t(jr.tw) * uint32_t(jr.th);
    if (blockIdx.x >= tiles) return;
    const draw_rec &d = f.draws[jr.draw];
    const brush_rec &br = f.brushes[d.brush];
    const cov_source cs = make_cov_source(f, sb);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *plane = f.planes + jr.plane_offset;
    for (uint32_t tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
        int tx = jr.tx0 + int(tl % uint32_t(jr.tw)), ty = jr.ty0 + int(tl / uint32_t(jr.tw));
        uint32_t te = jr.te_base + tl;
        int x = tx * kTile + lane;
        for (int k = 0; k < kTile / (kBlock / 32); ++k) {
            int ly = warp + k * (kBlock / 32), y = ty * kTile + ly;
            float sum = tile_row_sum(cs, te, ly, j, y, tx * kTile, row_buf[warp]);
            float cov = fminf(fabsf(sum), 1.0f);
            if (x < jr.left || x >= jr.left + jr.bw || y < jr.top || y >= jr.top + jr.bh) continue;
            float v = 0.0f;
            if (cov >= kThreshold) {
                vec2 centre = v2(float(x) + 0.5f, float(y) + 0.5f) - v2(jr.off_x, jr.off_y);
                v = cov * paint_alpha(f, br, d.inverse, centre);
            }
            plane[size_t(y - jr.top) * size_t(jr.bw) + size_t(x - jr.left)] = v;
        }
    }
}

// grid: (element stride, shadow job); axis 0 = along rows, 1 = along columns
__global__ void __launch_bounds__(kBlock) k_blur_pass(device_frame f, const float *src_base, float *dst_base,
                                                       int axis)
{
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const job_rec &jr = f.jobs[f.shadow_jobs[blockIdx.y]];
    const size_t n = size_t(jr.bw) * size_t(jr.bh);
    const float *src = src_base + jr.plane_offset;
    float *dst = dst_base + jr.plane_offset;
    const int r = jr.radius, bw = jr.bw, bh = jr.bh;
    const float w1 = jr.w1, w12 = jr.w1 + jr.w2;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        int x = int(i % size_t(bw)), y = int(i / size_t(bw));
        int at = axis ? y : x, len = axis ? bh : bw;
        size_t step = axis ? size_t(bw) : 1;
        const float *line = src + (axis ? size_t(x) : size_t(y) * size_t(bw));
        float inner = 0.0f;
        int lo = max(at - r, 0), hi = min(at + r, len - 1);
        for (int k = lo; k <= hi; ++k) inner += line[size_t(k) * step];
        float outer = 0.0f;
        if (at - r - 1 >= 0) outer += line[size_t(at - r - 1) * step];
        if (at + r + 1 < len) outer += line[size_t(at + r + 1) * step];
        dst[i] = w12 * inner + w1 * outer;
    }
}

}  // namespace

void launch_shadow(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    (void)t;
    if (!f.n_shadow_jobs) return;
    dim3 grid(64, f.n_shadow_jobs);
    k_shadow_raster<<<grid, kBlock, 0, s>>>(f, sorted_buffer);
    dim3 bgrid(128, f.n_shadow_jobs);
    const float *src = f.planes;
    float *dst = f.planes_tmp;
    for (int pass = 0; pass < 6; ++pass) {
        k_blur_pass<<<bgrid, kBlock, 0, s>>>(f, src, dst, pass >= 3 ? 1 : 0);
        const float *was = src;
        src = dst;
        dst = const_cast<float *>(was);
    }
}

}  // namespace cb200
