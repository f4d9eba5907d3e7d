// shadow.cu -- K9: drop shadows (reference render_shadow, hpp:2395-2539).
//
//   k_blur_x         x sweep FUSED with the shadow's alpha raster (hpp:2430-2452): a warp owns one tile row
//   k_blur_y         of the padded raster space and produces its samples as it sweeps -- one value per
//                    scanline in tiles without edges, the scanline's own runs otherwise -- so the alpha
//                    plane is never written or read.  Both sweeps run the three extended-box passes
//                    (Gwosdek et al.) of their axis, zero outside the working rectangle (hpp:2453-2503),
//                    as ONE streaming pass: a thread owns one line (a plane row for x, a storage column
//                    for y) and pushes every sample through three cascaded running sums whose 2r+3-deep
//                    histories live in shared memory.  A step is the reference's four-term update of the
//                    window sum (hpp:2463-2479) written as two differences and two fused multiply-adds;
//                    sweeps restart every chunk, so the result stays within ~1e-7 of the reference's own
//                    running sum (the plane is only ever a factor of the composite, hpp:2519-2523).
//   k_blur_units     prefix sums of the sweeps' work units (32 lines x one chunk) over all planes
//   k_shadow_raster  only for radii whose history does not fit (r > kStreamMaxRadius): coverage *
//   k_blur_rows      paint.alpha -> float plane per 32x32 tile (tile_cov.cuh), then the three passes per
//   k_transpose      row in shared memory, each output a windowed sum, with 32x32 tiled transposes
//                    around the column passes.
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

constexpr int kStreamThreads = 128;        // lines in flight per CTA; every warp works on its own
constexpr int kStreamMaxRadius = 30;       // history 3 x (2r+3) x 128 floats <= 95 KB
constexpr int kStreamChunk = 256;          // longer lines are swept in pieces (with a 3(r+1) run-in)

// grid: (tile stride, shadow job)
__global__ void __launch_bounds__(kBlock) k_shadow_raster(device_frame f, int sb)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t j = f.shadow_jobs[blockIdx.y];
    const job_rec &jr = f.jobs[j];
    const uint32_t tiles = uint32_t(jr.tw) * uint32_t(jr.th);
    if (jr.radius <= kStreamMaxRadius) return;                 // the x sweep rasters those itself (k_blur_x)
    if (blockIdx.x * (kBlock / 32) >= tiles) return;
    const draw_rec &d = f.draws[jr.draw];
    const brush_rec &br = f.brushes[d.brush];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *plane = f.planes + jr.plane_offset;
    // a solid brush has one alpha for the whole plane (negative: evaluate the brush per pixel)
    const float flat_alpha = br.type == CB200_BRUSH_COLOR ? (br.n_colors ? f.colors[br.first_color].w : 0.0f) : -1.0f;
    // one warp per tile: the tile set-up is paid once for 32 rows
    for (uint32_t tl = blockIdx.x * (kBlock / 32) + warp; tl < tiles; tl += gridDim.x * (kBlock / 32)) {
        int tx = jr.tx0 + int(tl % uint32_t(jr.tw)), ty = jr.ty0 + int(tl / uint32_t(jr.tw));
        uint32_t te = jr.te_base + tl;
        int x = tx * kTile + lane;
        const bool x_in = x >= jr.left && x < jr.left + jr.bw;
        // row info of the whole tile in three coalesced loads (lane = row); most rows have no run in the tile
        const float carried = f.te_backdrop[te * kTile + lane];
        const uint32_t first = f.te_first[te * kTile + lane];
        const uint32_t pixels = f.te_mask[te * kTile + lane];
        float *out = plane + ptrdiff_t(ty * kTile - jr.top) * ptrdiff_t(jr.pitch) + ptrdiff_t(x - jr.left + jr.skew);
        // rows of the plane this canvas (band) needs: [need_r0, need_r1) (job_rec)
        const int ly0 = max(0, jr.top + jr.need_r0 - ty * kTile), ly1 = min(kTile, jr.top + jr.need_r1 - ty * kTile);
        if (ly1 <= ly0) continue;
        // Most tiles of a shadow plane hold no edge at all (empty border, solid interior): every row is
        // one value, and the tile goes out as 16-byte stores, four rows per step (tile columns start on
        // a 128 B line: job_rec::skew).  Same arithmetic per row as below.
        if (flat_alpha >= 0.0f && !__ballot_sync(0xffffffffu, first != kNoRun) &&
            tx * kTile >= jr.left && tx * kTile + kTile <= jr.left + jr.bw) {
            const float row_cov = fminf(fabsf(carried), 1.0f);
            const float row_v = row_cov >= kThreshold ? row_cov * flat_alpha : 0.0f;     // lane = row
            float *tile_out = out - lane;                                                   // the tile's first column
            if (((tx * kTile - jr.left + jr.skew) & 3) == 0) {                              // 16-byte aligned (border % 4 == 0)
#pragma unroll
                for (int i = 0; i < kTile / 4; ++i) {
                    const int ly = i * 4 + (lane >> 3);
                    const float v = __shfl_sync(0xffffffffu, row_v, ly);
                    if (ly >= ly0 && ly < ly1)
                        reinterpret_cast<float4 *>(tile_out + ptrdiff_t(ly) * ptrdiff_t(jr.pitch))[lane & 7] = make_float4(v, v, v, v);
                }
            } else {
                for (int ly = ly0; ly < ly1; ++ly) out[ptrdiff_t(ly) * ptrdiff_t(jr.pitch)] = __shfl_sync(0xffffffffu, row_v, ly);
            }
            continue;
        }
        for (int ly = ly0; ly < ly1; ++ly) {
            const int y = ty * kTile + ly;
            const float sum = pixel_sum<true>(f.cumulative, __shfl_sync(0xffffffffu, carried, ly), __shfl_sync(0xffffffffu, first, ly),
                                              __shfl_sync(0xffffffffu, pixels, ly));
            float cov = fminf(fabsf(sum), 1.0f);
            float v = 0.0f;
            if (cov >= kThreshold) {
                if (flat_alpha >= 0.0f) v = cov * flat_alpha;
                else {
                    vec2 centre = v2(float(x) + 0.5f, float(y) + 0.5f) - v2(jr.off_x, jr.off_y);
                    v = cov * paint_alpha(f, br, d.inverse, centre);
                }
            }
            if (x_in) out[ptrdiff_t(ly) * ptrdiff_t(jr.pitch)] = v;
        }
    }
}

// ---- streaming blur ---------------------------------------------------------------------------

// Histories: ring[(slot * 3 + pass) * kStreamThreads + tid], slot in [0, 2r+3): the three passes of
// one slot sit at constant offsets from one another, and the slot a step reads is the slot the next
// step overwrites, so a step advances ONE address (add, compare, select).
constexpr int kRingSlot = 3 * kStreamThreads;                      // floats per slot
constexpr int kRingPass = kStreamThreads;                          // floats between passes

// One pass of the cascade: `run` is the window sum after the newest sample, `prev` that sample,
// `old` the sample that left the window last step (= what this step's read was last step).
struct stream_pass { float run, prev, old; };

// hpp:2463-2479 per output: running -= w2 s[x-r-1]; running -= w1 s[x-r-2]; running += w2 s[x+r];
// running += w1 s[x+r+1].  Here as two differences and two fused multiply-adds (the same sum, other
// rounding: the shadow plane only enters the composite as a factor, hpp:2519-2523 -- no decision is
// taken on it -- and a sweep restarts every chunk, so the difference stays near 1e-7).
// v = s[x+r+1] enters, leaving = s[x-r-1] (read from the ring), ps.old = s[x-r-2].
__device__ __forceinline__ float pass_step(stream_pass &ps, float v, float leaving, float w1, float w2)
{
    const float d2 = ps.prev - leaving, d1 = v - ps.old;
    ps.run = __fmaf_rn(w1, d1, __fmaf_rn(w2, d2, ps.run));
    ps.old = leaving;
    ps.prev = v;
    return ps.run;
}

// Prefix sums of every shadow job's sweep units (32 adjacent lines x one chunk), so that the
// sweeps can deal all units of a frame to warps as they become free -- plane sizes differ by
// orders of magnitude.  Also resets the two unit tickets.  One CTA.
__global__ void __launch_bounds__(kBlock) k_blur_units(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    if (f.hdr->overflow) return;
    const uint32_t n = f.n_shadow_jobs;
    uint32_t *along_x = f.blur_units, *along_y = f.blur_units + n + 1;
    uint32_t carry_x = 0, carry_y = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t ux = 0, uy = 0;
        if (i < n) {
            const job_rec &jr = f.jobs[f.shadow_jobs[i]];
            if (jr.radius <= kStreamMaxRadius && jr.need_r1 > jr.need_r0) {
                // x sweep: the tile rows (of the padded raster space) that hold rows [need_r0, need_r1), every chunk along the row;
                // y sweep: every 32-column strip, the chunks [chunk_lo, chunk_lo + chunk_n)
                ux = uint32_t((jr.top + jr.need_r1 - 1) / 32 - (jr.top + jr.need_r0) / 32 + 1) * uint32_t((jr.bw + kStreamChunk - 1) / kStreamChunk);
                uy = uint32_t((jr.pitch + 31) / 32) * uint32_t(jr.chunk_n);
            }
        }
        uint32_t total_x, total_y;
        const uint32_t ex = block_exclusive_scan(ux, sm, total_x);
        const uint32_t ey = block_exclusive_scan(uy, sm, total_y);
        if (i < n) { along_x[i] = carry_x + ex; along_y[i] = carry_y + ey; }
        carry_x += total_x; carry_y += total_y;
    }
    if (threadIdx.x == 0) {
        along_x[n] = carry_x; along_y[n] = carry_y;
        f.blur_units[2 * (n + 1)] = 0; f.blur_units[2 * (n + 1) + 1] = 0;
    }
}

__device__ __forceinline__ uint32_t shared_address(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
template <int kOffset>
__device__ __forceinline__ float ring_load(uint32_t at)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(at), "n"(kOffset));
    return v;
}
template <int kOffset>
__device__ __forceinline__ void ring_store(uint32_t at, float v)
{
    asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(at), "n"(kOffset), "f"(v));
}

// The three cascaded passes of one line.  Pass k's output x is final once its input x + r + 1 has
// entered, so sample t of the line yields output t - (r+1) of pass 1, t - 2(r+1) of pass 2 and
// t - 3(r+1) of pass 3.  The reference pads every pass with zeros outside the line: the outputs
// a pass produces left of sample 0 or right of sample len-1 enter the next pass as zero
// (kMasked; only the first 3(r+1) and the last 2(r+1) steps of a line need it).  Started from an
// all-zero state anywhere on a line the cascade is exact after 3(r+1) further samples: what entered
// a window before it was complete leaves it again with the same value.
struct cascade {
    stream_pass p1, p2, p3;
    float l1, l2, l3;                      // the samples leaving the windows at the next step
    uint32_t wr, rd, first, end;           // shared-memory byte addresses: slot written next, slot after it, ring bounds
    float w1, w2;
    int p, len;

    __device__ __forceinline__ void reset(float *ring, int W, int r, int line_len, float weight_1, float weight_2)
    {
        for (int k = 0; k < 3 * W; ++k) ring[k * kStreamThreads] = 0.0f;
        p1 = p2 = p3 = stream_pass{0.0f, 0.0f, 0.0f};
        l1 = l2 = l3 = 0.0f;
        first = shared_address(ring);
        end = first + uint32_t(W * kRingSlot) * 4u;
        wr = first; rd = first + kRingSlot * 4u;
        w1 = weight_1; w2 = weight_2; p = r + 1; len = line_len;
    }

    template <bool kMasked>
    __device__ __forceinline__ float push(float v, int t)
    {
        constexpr int kPassBytes = kRingPass * 4;
        ring_store<0>(wr, v);
        float o1 = pass_step(p1, v, l1, w1, w2);
        if (kMasked && unsigned(t - p) >= unsigned(len)) o1 = 0.0f;
        ring_store<kPassBytes>(wr, o1);
        float o2 = pass_step(p2, o1, l2, w1, w2);
        if (kMasked && unsigned(t - 2 * p) >= unsigned(len)) o2 = 0.0f;
        ring_store<2 * kPassBytes>(wr, o2);
        const float o3 = pass_step(p3, o2, l3, w1, w2);
        // the slot after next holds what leaves at the next step (2r+3 >= 3 slots: never one just written)
        uint32_t ahead = rd + kRingSlot * 4u;
        if (ahead == end) ahead = first;
        l1 = ring_load<0>(ahead); l2 = ring_load<kPassBytes>(ahead); l3 = ring_load<2 * kPassBytes>(ahead);
        wr = rd; rd = ahead;
        return o3;
    }
};

// The sweeps run as persistent grids of independent warps; each warp takes the next unit (32 adjacent lines x one
// chunk of some plane) off a ticket counter.  Returns false when the units are used up.
__device__ __forceinline__ bool next_unit(const device_frame &f, bool along_x, uint32_t &job_slot, int &unit)
{
    const uint32_t n_jobs = f.n_shadow_jobs;
    const uint32_t *prefix = f.blur_units + (along_x ? 0u : n_jobs + 1u);
    uint32_t *ticket = f.blur_units + 2 * (n_jobs + 1) + (along_x ? 0 : 1);
    uint32_t global_unit = 0;
    if ((threadIdx.x & 31) == 0) global_unit = atomicAdd(ticket, 1u);
    global_unit = __shfl_sync(0xffffffffu, global_unit, 0);
    if (global_unit >= prefix[n_jobs]) return false;
    uint32_t lo = 0, hi = n_jobs;                               // last job whose prefix <= global_unit
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (prefix[mid] <= global_unit) lo = mid; else hi = mid;
    }
    job_slot = lo;
    unit = int(global_unit - prefix[lo]);
    return true;
}

// x sweep, fused with the shadow's alpha raster (hpp:2430-2452): the lines are plane rows, a warp owns one tile row
// of the padded raster space (lane = scanline) and walks its 32x32 tiles left to right.  A tile's samples never
// exist in global memory: most tiles of a shadow plane hold no edge at all (empty border, solid interior), so a
// scanline's 32 samples are ONE value -- coverage carried in from the left times the paint's alpha -- that the lane
// already holds; tiles with edges are rastered like the compositor does it (lane = column, tile_cov.cuh) into a
// skewed shared tile and read back by row.  Outputs leave through the same tile, so global stores stay coalesced.
// Saves the raster's 4 B/pixel write and the sweep's 4 B/pixel read.
__global__ void __launch_bounds__(kStreamThreads, 5) k_blur_x(device_frame f, float *dst_base)
{
    grid_dependency_wait();
    extern __shared__ float blur_smem[];
    if (f.hdr->overflow) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *tile = blur_smem + warp * 32 * 33;
    float *mine = tile + lane * 33;
    float *ring = blur_smem + kStreamThreads * 33 + tid;
    uint32_t slot;
    int unit;
    while (next_unit(f, true, slot, unit)) {
        const job_rec &jr = f.jobs[f.shadow_jobs[slot]];
        const draw_rec &d = f.draws[jr.draw];
        const brush_rec &br = f.brushes[d.brush];
        const int r = jr.radius, W = 2 * r + 3, p = r + 1;
        const int len = jr.bw, pitch = jr.pitch;
        float *dst = dst_base + jr.plane_offset;
        // a solid brush has one alpha for the whole plane (negative: evaluate the brush per pixel)
        const float flat_alpha = br.type == CB200_BRUSH_COLOR ? (br.n_colors ? f.colors[br.first_color].w : 0.0f) : -1.0f;
        // tile rows that hold plane rows [need_r0, need_r1) (job_rec: what this canvas / band needs)
        const int ty_first = (jr.top + jr.need_r0) / 32;
        const int n_strips = (jr.top + jr.need_r1 - 1) / 32 - ty_first + 1;
        const int ty = ty_first + unit % n_strips, chunk = unit / n_strips;
        const int row = ty * 32 + lane - jr.top;                   // this lane's plane row
        const bool active = row >= jr.need_r0 && row < jr.need_r1;
        const int q_lo = max(jr.need_r0 - (ty * 32 - jr.top), 0), q_hi = min(jr.need_r1 - (ty * 32 - jr.top), 32);
        // outputs [y0, y1) of the line; they leave the cascade at steps [y0 + 3p, t_last)
        const int y0 = chunk * kStreamChunk, y1 = min(len, y0 + kStreamChunk);
        const int t_begin = chunk ? y0 - 3 * p : 0, t_last = y1 + 3 * p;
        const int plain_lo = 3 * p, plain_hi = len + p;            // steps in between need no masks
        cascade c;
        c.reset(ring, W, r, len, jr.w1, jr.w2);
        // blocks are tiles of the padded raster space: sample t of a line is padded x = left + t
        const int t_first = t_begin - ((t_begin + jr.left) & 31);
        float *out_rows = dst + ptrdiff_t(ty * 32 - jr.top) * ptrdiff_t(pitch) + ptrdiff_t(jr.skew + lane - 3 * p);
        // A tile's row info in three coalesced loads (lane = scanline), requested two tiles ahead; the running sums
        // after the scanline's first four runs inside the tile (tile_cov.cuh: `cumulative`, compacted per scanline)
        // one tile ahead, once the info that says where they are has arrived.
        struct tile_rows { float carried; uint32_t first, pixels; };
        struct tile_sums { float s0, s1, s2, s3; };
        auto request = [&](int tb) -> tile_rows {
            const int tx = (jr.left + tb) / 32;
            tile_rows ti = { 0.0f, kNoRun, 0u };
            if (tb < len && tx >= jr.tx0 && tx < jr.tx0 + jr.tw && ty >= jr.ty0 && ty < jr.ty0 + jr.th) {
                const uint32_t at = (jr.te_base + uint32_t(ty - jr.ty0) * uint32_t(jr.tw) + uint32_t(tx - jr.tx0)) * kTile + uint32_t(lane);
                ti.carried = f.te_backdrop[at]; ti.first = f.te_first[at];
                ti.pixels = ti.first != kNoRun ? f.te_mask[at] : 0u;
            }
            return ti;
        };
        auto request_sums = [&](const tile_rows &ti) -> tile_sums {
            tile_sums ts = { 0.0f, 0.0f, 0.0f, 0.0f };
            if (ti.first != kNoRun) {
                const float *at = f.cumulative + (ti.first & ~kRowEnds);
                const int n = __popc(ti.pixels);
                ts.s0 = at[0];                                     // first != kNoRun: at least one run
                if (n > 1) ts.s1 = at[1];
                if (n > 2) ts.s2 = at[2];
                if (n > 3) ts.s3 = at[3];
            }
            return ts;
        };
        tile_rows next_rows = request(t_first), after_rows = request(t_first + 32);
        tile_sums next_sums = request_sums(next_rows);
        for (int tb = t_first; tb < t_last; tb += 32) {
            const tile_rows rows = next_rows;
            tile_sums sums = next_sums;
            next_rows = after_rows;
            if (tb + 32 < t_last) next_sums = request_sums(next_rows);
            if (tb + 64 < t_last) after_rows = request(tb + 64);
            const float cr = rows.carried; const uint32_t fi = rows.first, pm = rows.pixels;
            const int x_tile = jr.left + tb;                        // padded x of the tile's first column
            const bool inside = tb >= 0 && tb + 32 <= len;          // every column of the tile is a sample of the line
            const bool edges = __ballot_sync(0xffffffffu, fi != kNoRun) != 0;
            const bool plain = tb >= plain_lo && tb + 32 <= plain_hi;
            // (warp-uniform: the last branch shuffles)
            if (flat_alpha >= 0.0f && !edges && (inside || __all_sync(0xffffffffu, cr == 0.0f))) {
                // one value per scanline (a tile that sticks out of the line qualifies only when it is empty)
                const float cov = fminf(fabsf(cr), 1.0f);
                const float v = cov >= kThreshold ? cov * flat_alpha : 0.0f;
                if (active) {
                    if (plain) {
#pragma unroll 8
                        for (int u = 0; u < 32; ++u) mine[u] = c.push<false>(v, 0);
                    } else {
#pragma unroll 8
                        for (int u = 0; u < 32; ++u) mine[u] = c.push<true>(unsigned(tb + u) < unsigned(len) ? v : 0.0f, tb + u);
                    }
                }
            } else if (flat_alpha >= 0.0f) {
                // edges: a lane walks ITS scanline's runs left to right -- the sum changes at the pixels of `pm`,
                // and after the scanline's last run nothing carries on (render_shadow walks the runs alone,
                // hpp:2430-2452; pixel_sum<true> says the same per pixel)
                if (active) {
                    const float *more = f.cumulative + (fi & ~kRowEnds);
                    const int last_pixel = (fi != kNoRun && (fi & kRowEnds)) ? 31 - __clz(int(pm)) : 32;
                    float sum = cr;
                    int taken = 0;
#pragma unroll 8
                    for (int u = 0; u < 32; ++u) {
                        if (pm >> u & 1u) {
                            if (taken < 4) { sum = sums.s0; sums.s0 = sums.s1; sums.s1 = sums.s2; sums.s2 = sums.s3; }
                            else sum = more[taken];
                            ++taken;
                        }
                        const float cov = u > last_pixel ? 0.0f : fminf(fabsf(sum), 1.0f);
                        float v = cov >= kThreshold ? cov * flat_alpha : 0.0f;
                        if (unsigned(tb + u) >= unsigned(len)) v = 0.0f;
                        mine[u] = c.push<true>(v, tb + u);
                    }
                }
            } else {
                // a brush that varies per pixel: raster the tile (lane = column, scanline by scanline), sweep it by row
                const int x = x_tile + lane;
                const bool x_in = tb + lane >= 0 && tb + lane < len;
#pragma unroll 1
                for (int ly = 0; ly < 32; ++ly) {
                    const float sum = pixel_sum<true>(f.cumulative, __shfl_sync(0xffffffffu, cr, ly), __shfl_sync(0xffffffffu, fi, ly),
                                                      __shfl_sync(0xffffffffu, pm, ly));
                    const float cov = fminf(fabsf(sum), 1.0f);
                    float v = 0.0f;
                    if (x_in && cov >= kThreshold && ly >= q_lo && ly < q_hi) {
                        const vec2 centre = v2(float(x) + 0.5f, float(ty * 32 + ly) + 0.5f) - v2(jr.off_x, jr.off_y);
                        v = cov * paint_alpha(f, br, d.inverse, centre);
                    }
                    tile[ly * 33 + lane] = v;
                }
                __syncwarp();
                if (active) {
#pragma unroll 8
                    for (int u = 0; u < 32; ++u) mine[u] = c.push<true>(mine[u], tb + u);
                }
            }
            __syncwarp();
            const int col = tb + lane - 3 * p;
            if (col >= y0 && col < y1) {
                float *out = out_rows + tb;
                if (q_lo == 0 && q_hi == 32) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) out[q * pitch] = tile[q * 33 + lane];
                } else {
                    for (int q = q_lo; q < q_hi; ++q) out[q * pitch] = tile[q * 33 + lane];
                }
            }
            __syncwarp();
        }
    }
}

// y sweep: the lines are storage columns, so that a warp always covers one aligned 128 B line of every row
__global__ void __launch_bounds__(kStreamThreads, 8) k_blur_y(device_frame f, const float *src_base, float *dst_base)
{
    grid_dependency_wait();
    extern __shared__ float blur_smem[];
    if (f.hdr->overflow) return;
    const int tid = threadIdx.x, lane = tid & 31;
    float *ring = blur_smem + tid;
    uint32_t slot;
    int unit;
    while (next_unit(f, false, slot, unit)) {
        const job_rec &jr = f.jobs[f.shadow_jobs[slot]];
        const int r = jr.radius, W = 2 * r + 3, p = r + 1;
        const int len = jr.bh, pitch = jr.pitch;
        const float *src = src_base + jr.plane_offset;
        float *dst = dst_base + jr.plane_offset;
        // every 32-column strip of the storage, the chunks [chunk_lo, chunk_lo + chunk_n) this canvas needs
        const int n_strips = (pitch + 31) / 32;
        const int strip = unit % n_strips, chunk = jr.chunk_lo + unit / n_strips;
        const int line = strip * 32 + lane;
        if (!(line >= jr.skew && line < jr.skew + jr.bw)) continue;
        // outputs [y0, y1) of the line; they leave the cascade at steps [t_store, t_last)
        const int y0 = chunk * kBlurChunkY, y1 = min(len, y0 + kBlurChunkY);
        const int t_begin = chunk ? y0 - 3 * p : 0, t_store = y0 + 3 * p, t_last = y1 + 3 * p;
        const int plain_lo = 3 * p, plain_hi = len + p;            // steps in between need no masks
        cascade c;
        c.reset(ring, W, r, len, jr.w1, jr.w2);
        constexpr int kAhead = 8;
        const float *in = src + size_t(line);
        float *out = dst + size_t(line);
        float ahead[kAhead];
        auto fetch = [&](int tb) {
            const float *from = in + ptrdiff_t(tb) * ptrdiff_t(pitch);
            // the rows three blocks further down are asked into L2 now: eight loads per lane in flight do not cover
            // HBM's latency at this bandwidth (52 % of the sweep's stall samples sat on these loads)
            if (tb + 4 * kAhead <= min(len, t_last)) {
                const float *later = from + 3 * kAhead * pitch;
#pragma unroll
                for (int u = 0; u < kAhead; ++u) asm volatile("prefetch.global.L2 [%0];" :: "l"(later + u * pitch));
            }
            if (tb >= 0 && tb + kAhead <= len) {
#pragma unroll
                for (int u = 0; u < kAhead; ++u) ahead[u] = from[u * pitch];
            } else {
#pragma unroll
                for (int u = 0; u < kAhead; ++u) ahead[u] = unsigned(tb + u) < unsigned(len) ? from[u * pitch] : 0.0f;
            }
        };
        // blocks of kAhead steps, laid out so that none straddles the first stored output
        const int t_first = t_store - ((t_store - t_begin + kAhead - 1) & ~(kAhead - 1));
        fetch(t_first);
        for (int tb = t_first; tb < t_last; tb += kAhead) {
            float v[kAhead];
#pragma unroll
            for (int u = 0; u < kAhead; ++u) v[u] = ahead[u];
            if (tb + kAhead < t_last) fetch(tb + kAhead);
            const bool plain = tb >= plain_lo && tb + kAhead <= plain_hi;
            if (plain && tb < t_store) {
#pragma unroll
                for (int u = 0; u < kAhead; ++u) c.push<false>(v[u], 0);
            } else if (plain && tb + kAhead <= t_last) {
                float *o = out + ptrdiff_t(tb - 3 * p) * ptrdiff_t(pitch);
#pragma unroll
                for (int u = 0; u < kAhead; ++u) o[u * pitch] = c.push<false>(v[u], 0);
            } else {
#pragma unroll
                for (int u = 0; u < kAhead; ++u) {
                    const float o = c.push<true>(v[u], tb + u);
                    const int i3 = tb + u - 3 * p;
                    if (i3 >= y0 && i3 < y1) out[size_t(i3) * size_t(pitch)] = o;
                }
            }
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
    grid_dependency_wait();
    extern __shared__ float blur_smem[];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const job_rec &jr = f.jobs[f.shadow_jobs[blockIdx.y]];
    const int len = transposed ? jr.bh : jr.bw, rows = transposed ? jr.bw : jr.bh;
    const int r = jr.radius, pad = r + 1;
    if (len <= 0 || rows <= 0 || r <= kStreamMaxRadius) return;
    const int stride = sk(len + 2 * pad) + 1;
    const float w1 = jr.w1, w12 = jr.w1 + jr.w2;
    // upright planes are pitched (job_rec::pitch, skew); transposed ones are compact bw rows of bh
    const size_t row_stride = transposed ? size_t(len) : size_t(jr.pitch);
    const float *src = src_base + jr.plane_offset + (transposed ? 0 : jr.skew);
    float *dst = dst_base + jr.plane_offset + (transposed ? 0 : jr.skew);
    constexpr int kWarpsPerCta = kBlock / 32;
    if (2 * stride * kWarpsPerCta <= smem_floats) {
        // short rows: every warp blurs its own row, eight rows per CTA in flight
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (int(blockIdx.x) * kWarpsPerCta >= rows) return;
        float *buf0 = blur_smem + size_t(warp) * 2 * stride, *buf1 = buf0 + stride;
        for (int i = lane; i < 2 * stride; i += 32) buf0[i] = 0.0f;
        __syncwarp();
        for (int row = blockIdx.x * kWarpsPerCta + warp; row < rows; row += gridDim.x * kWarpsPerCta)
            blur_one_row<32>(src + size_t(row) * row_stride, dst + size_t(row) * row_stride, len, r, w1, w12,
                             buf0, buf1, lane);
        return;
    }
    if (2 * stride > smem_floats || int(blockIdx.x) >= rows) return;      // host sized smem for the largest row
    float *buf0 = blur_smem, *buf1 = blur_smem + stride;
    for (int i = threadIdx.x; i < 2 * stride; i += kBlock) buf0[i] = 0.0f;
    __syncthreads();
    for (int row = blockIdx.x; row < rows; row += gridDim.x)
        blur_one_row<kBlock>(src + size_t(row) * row_stride, dst + size_t(row) * row_stride, len, r, w1, w12,
                             buf0, buf1, threadIdx.x);
}

// 32x32 tiled transpose of every shadow plane (bh x bw -> bw x bh or back), grid: (tile stride, job)
__global__ void __launch_bounds__(kBlock) k_transpose(device_frame f, const float *src_base, float *dst_base,
                                                       int back)
{
    grid_dependency_wait();
    __shared__ float tile[32][33];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const job_rec &jr = f.jobs[f.shadow_jobs[blockIdx.y]];
    const int sw = back ? jr.bh : jr.bw, sh = back ? jr.bw : jr.bh;   // source is sh rows of sw
    if (sw <= 0 || sh <= 0 || jr.radius <= kStreamMaxRadius) return;
    // the upright side is pitched, the transposed side compact
    const size_t src_stride = back ? size_t(sw) : size_t(jr.pitch), dst_stride = back ? size_t(jr.pitch) : size_t(sh);
    const float *src = src_base + jr.plane_offset + (back ? 0 : jr.skew);
    float *dst = dst_base + jr.plane_offset + (back ? jr.skew : 0);
    const int tiles_x = (sw + 31) / 32, tiles_y = (sh + 31) / 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;             // 32 x 8 threads
    for (int tl = blockIdx.x; tl < tiles_x * tiles_y; tl += gridDim.x) {
        const int x0 = (tl % tiles_x) * 32, y0 = (tl / tiles_x) * 32;
        for (int k = ly; k < 32; k += 8) {
            int x = x0 + lx, y = y0 + k;
            tile[k][lx] = (x < sw && y < sh) ? src[size_t(y) * src_stride + size_t(x)] : 0.0f;
        }
        __syncthreads();
        for (int k = ly; k < 32; k += 8) {
            int x = y0 + lx, y = x0 + k;                                  // destination is sw rows of sh
            if (x < sh && y < sw) dst[size_t(y) * dst_stride + size_t(x)] = tile[lx][k];
        }
        __syncthreads();
    }
}

}  // namespace

void launch_shadow(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s,
                   cudaEvent_t after_raster)
{
    if (!f.n_shadow_jobs) { if (after_raster) cudaEventRecord(after_raster, s); return; }
    // planes wider than the streaming sweeps take (r > kStreamMaxRadius) are rastered into memory first
    if (f.max_shadow_radius > kStreamMaxRadius) {
        dim3 grid(64, f.n_shadow_jobs);
        launch_pdl(k_shadow_raster, grid, kBlock, 0, s, f, sorted_buffer);
    }
    if (after_raster) cudaEventRecord(after_raster, s);
    const int longest = std::max(t.width, t.height) + f.max_shadow_pad;
    if (f.min_shadow_radius <= kStreamMaxRadius) {
        // x sweep planes -> planes_tmp, y sweep back into planes
        const int w = 2 * std::min(f.max_shadow_radius, kStreamMaxRadius) + 3;
        const size_t ring_bytes = size_t(3 * w * kStreamThreads) * sizeof(float);
        const size_t row_bytes = ring_bytes + size_t(kStreamThreads * 33) * sizeof(float);
        if (row_bytes > 48 * 1024) {                             // per device, so not cached in a static
            cudaFuncSetAttribute(k_blur_x, cudaFuncAttributeMaxDynamicSharedMemorySize, int(row_bytes));
            cudaFuncSetAttribute(k_blur_y, cudaFuncAttributeMaxDynamicSharedMemorySize, int(row_bytes));
        }
        auto resident = [](size_t smem) { return int(std::min<size_t>(16, (227 * 1024) / (smem + 1024))); };
        launch_pdl(k_blur_units, 1, kBlock, 0, s, f);
        // x sweep (rasters as it goes) -> planes_tmp, then the y sweep back into planes
        launch_pdl(k_blur_x, kSMs * resident(row_bytes), kStreamThreads, row_bytes, s, f, f.planes_tmp);
        launch_pdl(k_blur_y, kSMs * resident(ring_bytes), kStreamThreads, ring_bytes, s, f, f.planes_tmp, f.planes);
    }
    if (f.max_shadow_radius > kStreamMaxRadius) {
        // rows -> transpose -> rows (= columns) -> transpose back; the result ends up in f.planes
        // room for one longest row (CTA mode) and for eight rows of up to 512 pixels (warp mode)
        const int pad2 = 2 * (f.max_shadow_radius + 1);
        auto skewed = [](int n) { return n + (n >> 5) + 1; };
        const int smem_floats = std::max(2 * skewed(longest + pad2), 16 * skewed(std::min(longest, 512) + pad2));
        const size_t smem_bytes = size_t(smem_floats) * sizeof(float);
        if (smem_bytes > 48 * 1024)
            cudaFuncSetAttribute(k_blur_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
        dim3 rgrid(256, f.n_shadow_jobs), tgrid(128, f.n_shadow_jobs);
        launch_pdl(k_blur_rows, rgrid, kBlock, smem_bytes, s, f, f.planes, f.planes_tmp, 0, smem_floats);
        launch_pdl(k_transpose, tgrid, kBlock, 0, s, f, f.planes_tmp, f.planes, 0);
        launch_pdl(k_blur_rows, rgrid, kBlock, smem_bytes, s, f, f.planes, f.planes_tmp, 1, smem_floats);
        launch_pdl(k_transpose, tgrid, kBlock, 0, s, f, f.planes_tmp, f.planes, 1);
    }
}

}  // namespace cb200
