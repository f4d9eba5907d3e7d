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
constexpr int kBatch = kBlock / 32;                         // jobs staged in shared memory at a time

// Coverage of one tile row from staged row info (see tile_cov.cuh for the global
// memory variant used by the shadow rasteriser).
__device__ __forceinline__ float staged_row_sum(const cov_source &c, float backdrop, uint32_t first, uint32_t job,
                                                int y, int x0, float *row_buf)
{
    if (first == kNoRun) return backdrop;
    const int lane = threadIdx.x & 31;
    const uint64_t row_key = (uint64_t(job) << c.by) | uint64_t(uint32_t(y));
    const uint64_t xmask = (1ull << c.bx) - 1;
    row_buf[lane] = __int_as_float(0x7fc00000);
    __syncwarp();
    for (uint32_t k = first;; k += 32) {
        uint32_t idx = k + uint32_t(lane);
        bool ok = idx < c.n_runs;
        uint64_t key = ok ? c.keys[idx] : ~0ull;
        int col = int(key & xmask) - x0;
        ok = ok && (key >> c.bx) == row_key && col < kTile;
        if (ok) {
            float v = c.cumulative[idx];
            if (v == v) row_buf[col] = v;
        }
        if (!__all_sync(0xffffffffu, ok)) break;
    }
    __syncwarp();
    float mine = row_buf[lane];
    uint32_t have = __ballot_sync(0xffffffffu, mine == mine);
    uint32_t upto = have & (0xffffffffu >> (31 - lane));
    int src = upto ? 31 - __clz(upto) : 0;
    float got = __shfl_sync(0xffffffffu, mine, src);
    __syncwarp();
    return upto ? got : backdrop;
}

// Non-solid brushes: kept out of line so the solid path stays small.
__device__ __noinline__ rgba paint_slow(paint_tables f, uint32_t brush, uint32_t draw, float x, float y)
{
    return paint_at(f, f.brushes[brush], f.draws[draw].inverse, v2(x, y));
}

__global__ void __launch_bounds__(kBlock, 3) k_composite(device_frame f, canvas_target t, int sb,
                                                          int tiles_x, int tile_y0)
{
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_job[kBlock], s_te[kBlock];
    __shared__ int s_start;
    __shared__ __align__(16) comp_rec s_rec[kBatch];
    __shared__ float s_back[kBatch][kTile];
    __shared__ uint32_t s_first[kBatch][kTile];
    __shared__ float row_buf[kBlock / 32][kTile];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x % tiles_x, ty = tile_y0 + blockIdx.x / tiles_x;
    const int x = tx * kTile + lane;
    const int band_y1 = t.band_y0 + t.band_rows;
    const bool x_in = x < t.width;
    const int tile_x0 = tx * kTile, tile_y0p = ty * kTile;
    const uint32_t n_jobs = h->n_jobs;

    // Jobs are found through a compact 12-byte-per-job table (tile box + kind/flags word, first tile
    // entry): one coalesced 8 B load per job and thread instead of the 128 B record.
    auto box_hits = [&](uint2 box) -> bool {
        int bx0 = int(box.x & 0x7ffu), by0 = int((box.x >> 11) & 0x7ffu);
        int bx1 = int((box.x >> 22) & 0x3ffu) | int((box.y & 1u) << 10), by1 = int((box.y >> 1) & 0x7ffu);
        return tx >= bx0 && tx <= bx1 && ty >= by0 && ty <= by1;
    };
    auto entry_of = [&](uint2 box, uint32_t te_base) -> uint32_t {
        int bx0 = int(box.x & 0x7ffu), by0 = int((box.x >> 11) & 0x7ffu);
        int bx1 = int((box.x >> 22) & 0x3ffu) | int((box.y & 1u) << 10);
        return te_base + uint32_t(ty - by0) * uint32_t(bx1 - bx0 + 1) + uint32_t(tx - bx0);
    };

    // Pass 1 -- occlusion culling: the last job that paints this whole tile with an
    // opaque solid colour (covered tile entry, source_over/copy, alpha 1, unclipped)
    // makes every earlier job, and the old framebuffer content, irrelevant.
    if (threadIdx.x == 0) s_start = -1;
    __syncthreads();
    for (uint32_t base = 0; base < n_jobs; base += kBlock) {
        uint32_t j = base + threadIdx.x;
        if (j < n_jobs) {
            uint2 box = f.job_box[j];
            if ((box.y & JOBBOX_OPAQUE) && box_hits(box) && (f.te_flags[entry_of(box, f.job_te[j])] & TE_COVERED))
                atomicMax(&s_start, int(j));
        }
    }
    __syncthreads();
    const int start_job = s_start;

    float4 px[kRowsPerThread];
    int py[kRowsPerThread];
    bool live[kRowsPerThread];
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
        py[k] = ty * kTile + warp + k * (kBlock / 32);
        live[k] = x_in && py[k] >= t.band_y0 && py[k] < band_y1;
        px[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    bool loaded = start_job >= 0;              // a covering job replaces the old pixels: nothing to load
    bool touched = false;
    const cov_source cs = make_cov_source(f, sb);
    const paint_tables tables = { f.colors, f.stops, f.texels, f.brushes, f.draws };
    unsigned long long painted = 0;

    // Pass 2 -- replay the surviving jobs in submission order.
    const uint32_t first_job = start_job < 0 ? 0u : uint32_t(start_job);
    for (uint32_t base = first_job - first_job % kBlock; base < n_jobs; base += kBlock) {
        // which of these 256 jobs touch this tile?  (ordered compaction: ballot + per-warp counts)
        uint32_t j = base + threadIdx.x, te = 0;
        bool hit = false;
        if (j < n_jobs && j >= first_job) {
            uint2 box = f.job_box[j];
            if (box_hits(box)) {
                if ((box.y >> 12 & 3u) == JOB_SHADOW) hit = true;
                else {
                    te = entry_of(box, f.job_te[j]);
                    hit = (box.y & JOBBOX_EVERYWHERE) || (f.te_flags[te] & TE_NONEMPTY);
                }
            }
        }
        uint32_t votes = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) sm[warp] = __popc(votes);
        __syncthreads();
        uint32_t slot = __popc(votes & ((1u << lane) - 1u)), n_hit = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) {
            uint32_t c = sm[w];
            if (w < warp) slot += c;
            n_hit += c;
        }
        if (hit) { s_job[slot] = j; s_te[slot] = te; }
        __syncthreads();
        if (n_hit && !loaded) {                   // first job that touches the tile: fetch the old pixels
#pragma unroll
            for (int k = 0; k < kRowsPerThread; ++k)
                if (live[k]) px[k] = t.fb[size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x)];
            loaded = true;
        }
        touched = touched || n_hit != 0;

        for (uint32_t b0 = 0; b0 < n_hit; b0 += kBatch) {
            // stage up to kBatch jobs: one warp per job fetches its record and the 32
            // (backdrop, first run) pairs of this tile entry -- all loads in flight at once
            if (b0 + warp < n_hit) {
                uint32_t jj = s_job[b0 + warp], tte = s_te[b0 + warp];
                uint32_t word = reinterpret_cast<const uint32_t *>(&f.comp[jj])[lane];
                reinterpret_cast<uint32_t *>(&s_rec[warp])[lane] = word;
                bool has_rows = __shfl_sync(0xffffffffu, word, 0) != JOB_SHADOW;     // word 0 = kind
                s_back[warp][lane] = has_rows ? f.te_backdrop[tte * kTile + lane] : 0.0f;
                s_first[warp][lane] = has_rows ? f.te_first[tte * kTile + lane] : kNoRun;
            }
            __syncthreads();
            const uint32_t n_here = min(uint32_t(kBatch), n_hit - b0);
            for (uint32_t q = 0; q < n_here; ++q) {
                const comp_rec &c = s_rec[q];
                const uint32_t jj = s_job[b0 + q];
                const float *mask = c.mask_src ? t.mask_planes[c.mask_src] : nullptr;
                const uint32_t op = c.op;
                if (c.kind == JOB_SHADOW) {
                    const float *plane = f.planes + (uint64_t(c.plane_hi) << 32 | c.plane_lo);
                    const rgba tint = mk(c.color[0], c.color[1], c.color[2], c.color[3]);
#pragma unroll
                    for (int k = 0; k < kRowsPerThread; ++k) {
                        if (!live[k] || x < c.cx0 || x >= c.cx1 || py[k] < c.cy0 || py[k] >= c.cy1) continue;
                        float vis = mask ? fminf(fabsf(mask[size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x)]), 1.0f) : 1.0f;
                        if (vis < kThreshold) continue;
                        float s = plane[size_t(py[k] + c.border - c.top) * size_t(c.bw) + size_t(x + c.border - c.left)];
                        blend(px[k], scale(c.alpha * s, tint), op, vis);
                        ++painted;
                    }
                    continue;
                }
                const bool everywhere = (~op & 8u) != 0;
                const bool solid = c.brush_type == CB200_BRUSH_COLOR;
                const rgba flat = mk(c.color[0], c.color[1], c.color[2], c.color[3]);
                const float alpha = c.alpha;
                float *mask_out = c.kind == JOB_CLIP ? t.mask_planes[c.mask_dst] : nullptr;
#pragma unroll
                for (int k = 0; k < kRowsPerThread; ++k) {
                    const int ly = warp + k * (kBlock / 32);
                    // warp-uniform: every lane of the warp shares the row
                    float sum = staged_row_sum(cs, s_back[q][ly], s_first[q][ly], jj, py[k], tile_x0, row_buf[warp]);
                    float cov = fminf(fabsf(sum), 1.0f);
                    if (!live[k]) continue;
                    size_t at = size_t(py[k] - t.band_y0) * size_t(t.width) + size_t(x);
                    float vis = mask ? fminf(fabsf(mask[at]), 1.0f) : 1.0f;
                    if (mask_out) { mask_out[at] = cov * vis; continue; }
                    if (!((cov >= kThreshold || everywhere) && vis >= kThreshold)) continue;
                    rgba paint = solid ? flat
                                       : (c.brush_type == 0xffu ? mk(0.0f, 0.0f, 0.0f, 0.0f)
                                                                : paint_slow(tables, c.brush, c.draw, float(x) + 0.5f, float(py[k]) + 0.5f));
                    blend(px[k], scale(cov * alpha, paint), op, vis);
                    ++painted;
                }
            }
            __syncthreads();
        }
    }
    if (!touched) return;                          // no job reaches this tile: its pixels stay as they are
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
