// shadow.cu -- K9: drop shadows (reference render_shadow, hpp:2395-2539).
//
//   k_shadow_raster  per 32x32 tile of a shadow job's working rectangle (padded
//                    canvas space): coverage * paint.alpha -> float plane
//                    (hpp:2430-2452), coverage rebuilt from the sorted runs exactly
//                    like the compositor does (tile_cov.cuh)
//   k_blur_rows      the three extended-box passes (Gwosdek et al.) of one axis fused in shared
//                    memory, zero outside the working rectangle (hpp:2453-2503).  Each output is
//                      (w1+w2) * sum_{|d|<=r} s[i+d] + w1 * (s[i-r-1] + s[i+r+1])
//                    which is what the reference's running sum maintains.
//   k_transpose      32x32 tiled transpose so the column passes reuse the row kernel with
//                    coalesced accesses: rows, transpose, rows, transpose back.
// The blurred plane is consumed by the tile compositor (composite.cu).
#include "frame.cuh"
#include "tile_cov.cuh"

#include <algorithm>

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

// Three extended-box passes along the rows of a plane, fused: a CTA stages one row in shared
// memory (zero-padded by r + 1 on both sides, which is exactly the reference's "zero outside the
// working rectangle"), runs the three passes between two shared buffers and writes the row back --
// one global read and one global write per pixel for all three passes.  Each thread produces runs
// of 8 consecutive outputs with a sliding window (direct sum for the first, +new -old for the next
// seven), so shared-memory traffic is ~3 loads per output per pass.
// grid: (row stride, shadow job).  `transposed`: the plane is stored bw-major (after k_transpose).
constexpr int kBlurRun = 8;

// Shared-memory rows are skewed by one word per 32 so that threads walking runs of 8 consecutive
// pixels (addresses 8c + d) hit 32 different banks instead of 4.
__device__ __forceinline__ int sk(int i) { return i + (i >> 5); }

// One row by a group of `G` threads (a warp for short rows, the whole CTA for long ones).
template <int G>
__device__ __forceinline__ void blur_one_row(const float *in, float *out, int len, int r, float w1, float w12,
                                             float *buf0, float *buf1, int rank)
{
    const int pad = r + 1;
    auto group_sync = [] { if (G == 32) __syncwarp(); else __syncthreads(); };
    for (int i = rank; i < len; i += G) buf0[sk(pad + i)] = in[i];
    group_sync();
    float *from = buf0, *to = buf1;
    const int runs = (len + kBlurRun - 1) / kBlurRun;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        for (int c = rank; c < runs; c += G) {
            const int at = pad + c * kBlurRun;                   // buffer index of this run's first pixel
            float inner = 0.0f;
            for (int d = -r; d <= r; ++d) inner += from[sk(at + d)];
#pragma unroll
            for (int k = 0; k < kBlurRun; ++k) {
                if (c * kBlurRun + k < len) {
                    float lead = from[sk(at + k + r + 1)];
                    to[sk(at + k)] = w12 * inner + w1 * (from[sk(at + k - r - 1)] + lead);
                    inner += lead - from[sk(at + k - r)];
                }
            }
        }
        group_sync();
        float *t = from; from = to; to = t;
    }
    for (int i = rank; i < len; i += G) out[i] = from[sk(pad + i)];
    group_sync();
}

__global__ void __launch_bounds__(kBlock) k_blur_rows(device_frame f, const float *src_base, float *dst_base,
                                                       int transposed, int smem_floats)
{
    extern __shared__ float blur_smem[];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const job_rec &jr = f.jobs[f.shadow_jobs[blockIdx.y]];
    const int len = transposed ? jr.bh : jr.bw, rows = transposed ? jr.bw : jr.bh;
    const int r = jr.radius, pad = r + 1;
    if (len <= 0 || rows <= 0) return;
    const int stride = sk(len + 2 * pad) + 1;
    const float w1 = jr.w1, w12 = jr.w1 + jr.w2;
    const float *src = src_base + jr.plane_offset;
    float *dst = dst_base + jr.plane_offset;
    constexpr int kWarpsPerCta = kBlock / 32;
    if (2 * stride * kWarpsPerCta <= smem_floats) {
        // short rows: every warp blurs its own row, eight rows per CTA in flight
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (int(blockIdx.x) * kWarpsPerCta >= rows) return;
        float *buf0 = blur_smem + size_t(warp) * 2 * stride, *buf1 = buf0 + stride;
        for (int i = lane; i < 2 * stride; i += 32) buf0[i] = 0.0f;
        __syncwarp();
        for (int row = blockIdx.x * kWarpsPerCta + warp; row < rows; row += gridDim.x * kWarpsPerCta)
            blur_one_row<32>(src + size_t(row) * size_t(len), dst + size_t(row) * size_t(len), len, r, w1, w12,
                             buf0, buf1, lane);
        return;
    }
    if (2 * stride > smem_floats || int(blockIdx.x) >= rows) return;      // host sized smem for the largest row
    float *buf0 = blur_smem, *buf1 = blur_smem + stride;
    for (int i = threadIdx.x; i < 2 * stride; i += kBlock) buf0[i] = 0.0f;
    __syncthreads();
    for (int row = blockIdx.x; row < rows; row += gridDim.x)
        blur_one_row<kBlock>(src + size_t(row) * size_t(len), dst + size_t(row) * size_t(len), len, r, w1, w12,
                             buf0, buf1, threadIdx.x);
}

// 32x32 tiled transpose of every shadow plane (bh x bw -> bw x bh or back), grid: (tile stride, job)
__global__ void __launch_bounds__(kBlock) k_transpose(device_frame f, const float *src_base, float *dst_base,
                                                       int back)
{
    __shared__ float tile[32][33];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const job_rec &jr = f.jobs[f.shadow_jobs[blockIdx.y]];
    const int sw = back ? jr.bh : jr.bw, sh = back ? jr.bw : jr.bh;   // source is sh rows of sw
    if (sw <= 0 || sh <= 0) return;
    const float *src = src_base + jr.plane_offset;
    float *dst = dst_base + jr.plane_offset;
    const int tiles_x = (sw + 31) / 32, tiles_y = (sh + 31) / 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;             // 32 x 8 threads
    for (int tl = blockIdx.x; tl < tiles_x * tiles_y; tl += gridDim.x) {
        const int x0 = (tl % tiles_x) * 32, y0 = (tl / tiles_x) * 32;
        for (int k = ly; k < 32; k += 8) {
            int x = x0 + lx, y = y0 + k;
            tile[k][lx] = (x < sw && y < sh) ? src[size_t(y) * size_t(sw) + size_t(x)] : 0.0f;
        }
        __syncthreads();
        for (int k = ly; k < 32; k += 8) {
            int x = y0 + lx, y = x0 + k;                                  // destination is sw rows of sh
            if (x < sh && y < sw) dst[size_t(y) * size_t(sh) + size_t(x)] = tile[lx][k];
        }
        __syncthreads();
    }
}

}  // namespace

void launch_shadow(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    if (!f.n_shadow_jobs) return;
    dim3 grid(64, f.n_shadow_jobs);
    k_shadow_raster<<<grid, kBlock, 0, s>>>(f, sorted_buffer);
    // rows -> transpose -> rows (= columns) -> transpose back; the result ends up in f.planes
    const int longest = std::max(t.width, t.height) + f.max_shadow_pad;
    // room for one longest row (CTA mode) and for eight rows of up to 512 pixels (warp mode)
    const int pad2 = 2 * (f.max_shadow_radius + 1);
    auto skewed = [](int n) { return n + (n >> 5) + 1; };
    const int smem_floats = std::max(2 * skewed(longest + pad2), 16 * skewed(std::min(longest, 512) + pad2));
    const size_t smem_bytes = size_t(smem_floats) * sizeof(float);
    static size_t configured = 0;
    if (smem_bytes > 48 * 1024 && smem_bytes > configured) {
        cudaFuncSetAttribute(k_blur_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
        configured = smem_bytes;
    }
    dim3 rgrid(256, f.n_shadow_jobs), tgrid(128, f.n_shadow_jobs);
    k_blur_rows<<<rgrid, kBlock, smem_bytes, s>>>(f, f.planes, f.planes_tmp, 0, smem_floats);
    k_transpose<<<tgrid, kBlock, 0, s>>>(f, f.planes_tmp, f.planes, 0);
    k_blur_rows<<<rgrid, kBlock, smem_bytes, s>>>(f, f.planes, f.planes_tmp, 1, smem_floats);
    k_transpose<<<tgrid, kBlock, 0, s>>>(f, f.planes_tmp, f.planes, 1);
}

}  // namespace cb200
