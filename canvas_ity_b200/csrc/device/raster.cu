// raster.cu -- K4: scan conversion of every loop edge of every job into
// signed-area pixel runs (reference add_runs hpp:2109-2170 and the clip/clamp
// part of lines_to_runs hpp:2193-2240).  sm_100a, compiled with -fmad=false.
//
// Decomposition (all counts stay on the device):
//   k_job_items   one CTA: K4 work items per job = loop points of its draw
//   k_edges       one thread per (job, loop point): edge prev->point, offset,
//                 clipped per edge against the padded canvas (per-edge
//                 clipping == Sutherland-Hodgman for area accumulation) into <= 3
//                 pieces; counts the scanlines each piece touches inside the band
//   k_scan_rows   piece scanline counts -> exclusive offsets
//   k_row_count   one thread per (piece, scanline): closed-form run count
//   k_row_emit    same items: walks the pixels of one scanline exactly like the
//                 reference DDA (its per-row / per-column positions are evaluated
//                 directly from the edge's `from`, so any row can start cold) and
//                 writes (key = job|y|x, delta) pairs
//   k_job_tiles   one CTA: bounding boxes -> tile rectangles, tile-entry bases,
//                 shadow working rectangles and plane storage (hpp:2409-2428)
#include "frame.cuh"
#include "edge_clip.cuh"

namespace cb200 {

namespace {

// ------------------------------------------------------------- job items ----

// The two per-job bookkeeping kernels run as ONE CTA (their prefix sums carry from job to job): 1024 threads, so a
// batch of 18 K jobs takes 18 rounds of it, not 72.
constexpr int kOneCta = 1024;

__global__ void __launch_bounds__(kOneCta) k_job_items(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    uint32_t n = h->n_jobs, carry = 0;
    uint32_t stroke_base = h->n_line_points + h->n_dash_points;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t j = base + threadIdx.x, count = 0, first_point = 0;
        if (j < n && !h->overflow) {
            const draw_rec &d = f.draws[f.jobs[j].draw];
            if (d.kind == CB200_STROKE) {
                uint2 src = f.draw_src[f.jobs[j].draw];
                if (src.y > src.x) {
                    uint32_t a = f.half_offset[2 * src.x], b = f.half_offset[2 * src.y];
                    first_point = stroke_base + a;
                    count = b - a;
                }
            } else if (d.n_units) {
                uint32_t a = f.unit_offset[d.first_unit], b = f.unit_offset[d.first_unit + d.n_units];
                first_point = a;
                count = b - a;
            }
        }
        uint32_t total;
        uint32_t ex = block_exclusive_scan(count, sm, total);
        if (j < n) {
            job_rec &jr = f.jobs[j];
            jr.first_point = first_point;
            jr.first_item = carry + ex;
            jr.n_items = count;
            jr.min_x = jr.min_y = 0x7fffffff;
            jr.max_x = jr.max_y = -1;
            jr.run_min_x = jr.run_min_y = 0x7fffffff;
            jr.run_max_x = jr.run_max_y = -1;
            jr.first_key = 0xffffffffu;
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        h->n_items = carry;
        if (carry > f.cap_items) atomicOr(&h->overflow, OVF_ITEMS);
    }
}

// ------------------------------------------------------------------ edges ----

// Work mapping of the four kernels below: a CTA owns a contiguous slice of the items in whole
// tiles of kBlock, and thread t takes item `tile + t` -- neighbouring lanes read and write
// neighbouring records, whatever the frame size (a batch frame has ~10^8 row items).  Count
// kernels leave one total per CTA (-> exclusive bases by finish_partials), emit kernels scan
// tile by tile on top of their CTA's base.
__device__ __forceinline__ void tile_slice(uint32_t n, uint32_t &begin, uint32_t &end)
{
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kBlock - 1) / kBlock * kBlock;
    uint64_t b = uint64_t(per) * blockIdx.x;
    begin = b < n ? uint32_t(b) : n;
    end = b + per < n ? uint32_t(b + per) : n;
}

// Job of item `it`; the lanes of a warp hold consecutive items and `j` is the lane's previous
// answer.  One binary search per warp (first lane's item), then a short walk.
__device__ __forceinline__ uint32_t job_of_item(const device_frame &f, uint32_t n_jobs, uint32_t it, bool valid, uint32_t j)
{
    const bool stale = valid && !(it >= f.jobs[j].first_item && it < f.jobs[j].first_item + f.jobs[j].n_items);
    if (!__any_sync(0xffffffffu, stale)) return j;
    const uint32_t it0 = __shfl_sync(0xffffffffu, it, 0);
    uint32_t lo = 0, hi = n_jobs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (f.jobs[mid].first_item <= it0) lo = mid; else hi = mid;
    }
    if (valid)
        while (it >= f.jobs[lo].first_item + f.jobs[lo].n_items) ++lo;
    return lo;
}

__global__ void __launch_bounds__(kBlock) k_edges(device_frame f, canvas_target t)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    uint32_t n = h->overflow ? 0 : h->n_items, begin, end;
    tile_slice(n, begin, end);
    const uint32_t n_jobs = h->n_jobs;
    uint32_t sum = 0, j = 0;
    for (uint32_t tile = begin; tile < end; tile += kBlock) {
        const uint32_t it = tile + threadIdx.x;
        const bool valid = it < end;
        int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -1, by1 = -1;     // pixel bounds of this item's pieces
        j = job_of_item(f, n_jobs, it, valid, j);
        if (valid) {
        const job_rec &job = f.jobs[j];
        uint32_t p = job.first_point + (it - job.first_item);
        loop_span loop = f.loops[f.pt_loop[p]];
        uint32_t q = p == loop.first ? loop.first + loop.count - 1 : p - 1;
        float2 pa = f.pts[q], pb = f.pts[p];
        vec2 a = v2(job.off_x + pa.x, job.off_y + pa.y), b = v2(job.off_x + pb.x, job.off_y + pb.y);
        const float w = float(t.width + job.pad), hgt = float(t.height + job.pad);
        // the reference's four clip stages applied to this edge alone (edge_clip.cuh): the surviving part has
        // the polygon clip's own end points, the parts cut away beside the canvas are projected onto its sides
        clipped_edge ce;
        clip_edge(a, b, w, hgt, ce);
        if (job.kind == JOB_SHADOW && ce.n_events) {
            // The reference's polygon clip joins every sideways exit of a loop to the next entry by a boundary
            // segment whose runs count in render_shadow's bounding box (hpp:2208-2229, 2409-2419); projected
            // pieces do not reproduce that (k_row_emit ignores them for the box).  A loop with any crossing of
            // a side or of the top is listed once; k_shadow_boxes walks its crossings in order.
            const uint32_t loop_id = f.pt_loop[p];
            if (atomicExch(&f.loop_mark[loop_id], 1u) == 0u)
                f.box_loops[atomicAdd(&h->n_box_loops, 1u)] = make_uint2(j, loop_id);
        }
        int row0 = job.kind == JOB_SHADOW ? 0 : t.band_y0;
        int row1 = job.kind == JOB_SHADOW ? t.height + job.pad : t.band_y0 + t.band_rows;
        int made = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i >= ce.n_pieces) continue;
            const float4 pc = ce.piece[i];
            uint32_t rows = 0, rlo = 0;
            if (piece_has_runs(pc, ce.projected[i] != 0)) {
                edge_walk e = edge_setup(pc);
                // scanlines inside [row0, row1)
                int r_first, r_last;                       // inclusive range of r
                if (e.down) { r_first = row0 - int(e.fy0); r_last = row1 - 1 - int(e.fy0); }
                else { r_first = int(e.fy0) - (row1 - 1); r_last = int(e.fy0) - row0; }
                if (r_first < 0) r_first = 0;
                if (r_last > e.rows - 1) r_last = e.rows - 1;
                if (r_last >= r_first) {
                    rows = uint32_t(r_last - r_first + 1);
                    rlo = uint32_t(r_first);
                    int y_lo = e.down ? int(e.fy0) + r_first : int(e.fy0) - r_last;
                    int y_hi = e.down ? int(e.fy0) + r_last : int(e.fy0) - r_first;
                    int x_lo = int(e.fx0), x_hi = int(floorf(e.to.x)) + 1;
                    bx0 = min(bx0, x_lo); bx1 = max(bx1, x_hi);
                    by0 = min(by0, y_lo); by1 = max(by1, y_hi);
                }
            }
            uint32_t slot = it * 3 + made;
            f.pieces[slot] = pc;
            // bit 31: the piece lies left / right of the padded canvas and was projected onto its boundary.
            // Its runs only cancel one another; the reference's clip replaces such an excursion by ONE
            // boundary segment between the two crossing points (hpp:2208-2229), so these runs must not widen
            // the bounding box render_shadow takes from the runs (hpp:2409-2419).
            f.piece_job[slot] = j | (ce.projected[i] ? 0x80000000u : 0u);
            f.piece_rows[slot] = rows;
            f.piece_rlo[slot] = rlo;
            sum += rows;
            ++made;
        }
        for (; made < 3; ++made) f.piece_rows[it * 3 + made] = 0;
        }
        // job bounding boxes: neighbouring items nearly always belong to the same job, so lanes
        // with equal job combine their bounds first and one of them issues the four atomics
        const int lane = threadIdx.x & 31;
        const bool has = valid && bx1 >= 0;
        uint32_t peers = __match_any_sync(0xffffffffu, has ? j : 0x80000000u | uint32_t(lane));
        int gx0 = __reduce_min_sync(peers, bx0), gy0 = __reduce_min_sync(peers, by0);
        int gx1 = __reduce_max_sync(peers, bx1), gy1 = __reduce_max_sync(peers, by1);
        if (has && lane == __ffs(int(peers)) - 1) {
            atomicMin(&f.jobs[j].min_x, gx0); atomicMax(&f.jobs[j].max_x, gx1);
            atomicMin(&f.jobs[j].min_y, gy0); atomicMax(&f.jobs[j].max_y, gy1);
        }
    }
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[4 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + 4 * kGrid, &h->tickets[4], &h->n_row_items, sm);
}

// piece_rows (counts) -> piece_row_off (exclusive offsets), same slice mapping
__global__ void __launch_bounds__(kBlock) k_scan_rows(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    uint32_t n = h->overflow ? 0 : h->n_items, begin, end;
    tile_slice(n, begin, end);
    uint32_t carry = f.partials[4 * kGrid + blockIdx.x];
    for (uint32_t tile = begin; tile < end; tile += kBlock) {
        const uint32_t it = tile + threadIdx.x, s = it * 3;
        uint32_t r0 = 0, r1 = 0, r2 = 0;
        if (it < end) { r0 = f.piece_rows[s]; r1 = f.piece_rows[s + 1]; r2 = f.piece_rows[s + 2]; }
        uint32_t total;
        const uint32_t at = carry + block_exclusive_scan(r0 + r1 + r2, sm, total);
        if (it < end) { f.piece_row_off[s] = at; f.piece_row_off[s + 1] = at + r0; f.piece_row_off[s + 2] = at + r0 + r1; }
        carry += total;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && h->n_row_items > f.cap_rows) atomicOr(&h->overflow, OVF_ROWS);
}

// Largest piece slot whose offset is <= item (slots without rows share their offset with the
// next slot that has some, so this is the slot that contains the item).
__device__ __forceinline__ uint32_t find_piece(const uint32_t *off, uint32_t n_slots, uint32_t item)
{
    uint32_t lo = 0, hi = n_slots;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// The same for a whole warp whose lanes hold consecutive items: one binary search (first lane's
// item), one coalesced load of the next 32 offsets, five shuffles per lane; only lanes whose slot
// lies beyond that window (long stretches of pieces without rows: band canvases) search again.
__device__ __forceinline__ uint32_t piece_of_row_item(const uint32_t *off, uint32_t n_slots, uint32_t it, bool valid)
{
    const int lane = threadIdx.x & 31;
    const uint32_t s0 = find_piece(off, n_slots, __shfl_sync(0xffffffffu, it, 0));
    const uint32_t mine = s0 + uint32_t(lane) < n_slots ? off[s0 + uint32_t(lane)] : 0xffffffffu;
    uint32_t k = 0;
#pragma unroll
    for (uint32_t step = 16; step; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, mine, int(k + step));
        if (v <= it) k += step;
    }
    uint32_t slot = s0 + k;
    if (valid && k == 31 && s0 + 32 < n_slots && off[s0 + 32] <= it) slot = find_piece(off, n_slots, it);
    return slot;
}

__global__ void __launch_bounds__(kBlock) k_row_count(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    uint32_t n = h->overflow ? 0 : h->n_row_items, begin, end;
    tile_slice(n, begin, end);
    const uint32_t n_slots = h->n_items * 3;
    uint32_t sum = 0;
    for (uint32_t tile = begin; tile < end; tile += kBlock) {
        const uint32_t it = tile + threadIdx.x;
        const bool valid = it < end;
        const uint32_t slot = piece_of_row_item(f.piece_row_off, n_slots, valid ? it : end - 1, valid);
        if (!valid) continue;
        edge_walk e = edge_setup(f.pieces[slot]);
        row_walk w = row_setup(e, int(f.piece_rlo[slot] + (it - f.piece_row_off[slot])));
        uint32_t runs = uint32_t(w.inner) + 2;
        f.row_runs[it] = runs;
        f.row_piece[it] = slot;
        sum += runs;
    }
    uint32_t total;
    block_exclusive_scan(sum, sm, total);
    if (threadIdx.x == 0) f.partials[5 * kGrid + blockIdx.x] = total;
    finish_partials(f.partials + 5 * kGrid, &h->tickets[5], &h->n_runs, sm);
}

// walk_row_runs sink: (key, delta) pairs of one scanline, and the columns of its non-zero deltas
struct run_sink {
    uint64_t *keys; float *vals; uint64_t row_key;
    int lo_x, hi_x;
    __device__ __forceinline__ void put(float px, float delta)
    {
        *keys++ = row_key | uint64_t(uint32_t(px));
        *vals++ = delta;
        if (delta != 0.0f) { lo_x = min(lo_x, int(px)); hi_x = max(hi_x, int(px)); }
    }
};

__global__ void __launch_bounds__(kBlock) k_row_emit(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    if (h->n_runs > f.cap_runs) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&h->overflow, OVF_RUNS);
        return;
    }
    uint32_t n = h->overflow ? 0 : h->n_row_items, begin, end;
    tile_slice(n, begin, end);
    uint32_t carry = f.partials[5 * kGrid + blockIdx.x];
    const uint32_t bx = h->sort_bits_x, by = h->sort_bits_y;
    uint64_t *keys = f.keys[0];
    float *vals = f.vals[0];
    for (uint32_t tile = begin; tile < end; tile += kBlock) {
        const uint32_t it = tile + threadIdx.x;
        const bool valid = it < end;
        uint32_t total;
        uint32_t at = carry + block_exclusive_scan(valid ? f.row_runs[it] : 0u, sm, total);
        carry += total;
        // (no early exit for the lanes past the end: the bounds below are combined across the warp)
        uint32_t j = 0;
        bool projected = false, shadow = false;
        int lo_x = 0x7fffffff, hi_x = -1, row_y = 0;
        uint32_t first_key = 0xffffffffu;
        if (valid) {
            const uint32_t slot = f.row_piece[it];
            const uint32_t tagged = f.piece_job[slot];
            j = tagged & 0x7fffffffu;
            projected = (tagged >> 31) != 0;                     // an excursion outside the canvas, flattened onto its boundary
            edge_walk e = edge_setup(f.pieces[slot]);
            row_walk w = row_setup(e, int(f.piece_rlo[slot] + (it - f.piece_row_off[slot])));
            run_sink sink = { keys + at, vals + at, ((uint64_t(j) << by) | uint64_t(uint32_t(w.py))) << bx, 0x7fffffff, -1 };
            walk_row_runs(e, w, sink);
            shadow = f.jobs[j].kind == JOB_SHADOW;
            lo_x = sink.lo_x; hi_x = sink.hi_x; row_y = int(w.py);
            first_key = (uint32_t(w.py) << 16) | uint32_t(w.px);
        }
        if (f.n_shadow_jobs) {
            // Shadow jobs: exact bounds of the runs the reference keeps (non-zero deltas) plus the
            // smallest (y,x) run of all, which it keeps unconditionally (hpp:2244-2252).  A projected
            // piece enters nothing: its runs only cancel one another, and the boundary segment the
            // reference's clip puts in its place is accounted for by k_shadow_boxes.  Neighbouring row
            // items nearly always belong to one job, so the lanes of a job combine their bounds first and
            // one of them issues the atomics (one set per row item kept 1.1 M row items queueing on 305
            // words: k_row_emit 147 us on config 3).
            const int lane = threadIdx.x & 31;
            const bool counts = shadow && !projected;
            const uint32_t peers = __match_any_sync(0xffffffffu, counts ? j : 0x80000000u | uint32_t(lane));
            const bool has_runs = counts && hi_x >= 0;
            const int g_lo_x = __reduce_min_sync(peers, has_runs ? lo_x : 0x7fffffff);
            const int g_hi_x = __reduce_max_sync(peers, has_runs ? hi_x : -1);
            const int g_lo_y = __reduce_min_sync(peers, has_runs ? row_y : 0x7fffffff);
            const int g_hi_y = __reduce_max_sync(peers, has_runs ? row_y : -1);
            const uint32_t g_first = __reduce_min_sync(peers, counts ? first_key : 0xffffffffu);
            if (counts && lane == __ffs(int(peers)) - 1) {
                job_rec &jr = f.jobs[j];
                if (g_hi_x >= 0) {
                    atomicMin(&jr.run_min_x, g_lo_x); atomicMax(&jr.run_max_x, g_hi_x);
                    atomicMin(&jr.run_min_y, g_lo_y); atomicMax(&jr.run_max_y, g_hi_y);
                }
                atomicMin(&jr.first_key, g_first);
            }
        }
    }
}

// ---------------------------------------------------------- shadow boxes ----

// One warp per listed loop (a loop of a shadow job that crosses a side or the top of its padded canvas):
// lanes take 32 consecutive edges, the lanes that found crossings feed them in edge order to the walk of
// shadow_box.cuh (its state is replicated in every lane), and lane 0 enters what the reference's boundary
// segments add to the job's run bounding box.
__global__ void __launch_bounds__(kBlock) k_shadow_boxes(device_frame f, canvas_target t)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t n = h->n_box_loops;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n; i += n_warps) {
        const uint2 item = f.box_loops[i];
        const job_rec &job = f.jobs[item.x];
        const loop_span loop = f.loops[item.y];
        const float w = float(t.width + job.pad), hgt = float(t.height + job.pad);
        shadow_box_walk walk;
        walk.init(w, hgt);
        for (uint32_t base = 0; base < loop.count; base += 32) {
            const uint32_t k = base + uint32_t(lane);
            clipped_edge ce;
            ce.n_events = 0;
            ce.ev[0].v = ce.ev[1].v = ce.ev[2].v = 0.0f; ce.ev[0].kind = ce.ev[1].kind = ce.ev[2].kind = 0;
            if (k < loop.count) {
                const uint32_t p = loop.first + k, q = k ? p - 1 : loop.first + loop.count - 1;
                const float2 pa = f.pts[q], pb = f.pts[p];
                clip_edge(v2(job.off_x + pa.x, job.off_y + pa.y), v2(job.off_x + pb.x, job.off_y + pb.y), w, hgt, ce);
            }
            const int nev = ce.n_events;
            uint32_t mask = __ballot_sync(0xffffffffu, nev > 0);
            while (mask) {
                const int src = __ffs(int(mask)) - 1;
                mask &= mask - 1;
                const int n_src = __shfl_sync(0xffffffffu, nev, src);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const float v = __shfl_sync(0xffffffffu, ce.ev[e].v, src);
                    const int kind = __shfl_sync(0xffffffffu, ce.ev[e].kind, src);
                    if (e < n_src) walk.consume(kind, v);
                }
            }
        }
        walk.finish();
        if (lane == 0 && walk.hx >= 0) {
            job_rec &jr = f.jobs[item.x];
            atomicMin(&jr.run_min_x, walk.lx); atomicMax(&jr.run_max_x, walk.hx);
            atomicMin(&jr.run_min_y, walk.ly); atomicMax(&jr.run_max_y, walk.hy);
        }
    }
}

// ------------------------------------------------------------- job tiles ----

__global__ void __launch_bounds__(kOneCta) k_job_tiles(device_frame f, canvas_target t)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    __shared__ unsigned long long plane_carry, working_carry;
    frame_header *h = f.hdr;
    if (threadIdx.x == 0) { plane_carry = 0; working_carry = 0; }
    __syncthreads();
    uint32_t n = h->n_jobs, carry = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t j = base + threadIdx.x, tiles = 0;
        unsigned long long plane = 0;
        if (j < n && !h->overflow) {
            job_rec &jr = f.jobs[j];
            const draw_rec &d = f.draws[jr.draw];
            int x0, y0, x1, y1;                                  // raster-space pixel rectangle
            bool everywhere = jr.kind == JOB_CLIP || (jr.kind == JOB_MAIN && (~d.op & 8u));
            if (jr.kind == JOB_SHADOW) {
                // working rectangle of render_shadow, hpp:2409-2425
                int b = jr.border, full_w = t.width + 2 * b, full_h = t.height + 2 * b;
                int lx = jr.run_min_x, hx = jr.run_max_x, ly = jr.run_min_y, hy = jr.run_max_y;
                if (jr.first_key != 0xffffffffu) {
                    int fy = int(jr.first_key >> 16), fx = int(jr.first_key & 0xffffu);
                    lx = min(lx, fx); hx = max(hx, fx); ly = min(ly, fy); hy = max(hy, fy);
                }
                // nothing kept at all (columns come from runs, rows from inside runs and sideways crossings)
                if (hx < 0 || hy < 0) { lx = full_w; hx = 0; ly = full_h; hy = 0; }
                int left = max(lx - b, 0), right = min(hx + b, full_w) + 1;
                int top = max(ly - b, 0), bottom = min(hy + b, full_h);
                jr.left = left; jr.top = top;
                jr.bw = max(right - left, 0); jr.bh = max(bottom - top, 0);
                jr.skew = (left - b) & 31;                        // (left - skew - border) = 0 mod 32, also when negative
                jr.pitch = (jr.skew + jr.bw + 31) & ~31;
                plane = (unsigned long long)jr.pitch * (unsigned long long)jr.bh;
                x0 = left; y0 = top; x1 = left + jr.bw; y1 = top + jr.bh;
                // where the blurred plane lands on the canvas, hpp:2512-2514, 2535
                jr.cx0 = max(left - b, 0); jr.cx1 = min(right - b, t.width);
                jr.cy0 = max(top - b, t.band_y0); jr.cy1 = min(bottom - b, t.band_y0 + t.band_rows);
                if (plane == 0) { jr.cx1 = jr.cx0; jr.cy1 = jr.cy0; }
                // the plane rows this band composites, the y-sweep chunks that hold them, the rows those read
                // (hpp:2402-2405: the blur couples rows over 3 (r + 1) = border rows at most)
                const int ya = min(max(jr.cy0 + b - top, 0), jr.bh), yb = min(max(jr.cy1 + b - top, 0), jr.bh);
                jr.need_r0 = jr.need_r1 = 0; jr.chunk_lo = 0; jr.chunk_n = 0;
                if (plane && yb > ya) {
                    jr.chunk_lo = ya / kBlurChunkY;
                    jr.chunk_n = (yb - 1) / kBlurChunkY - jr.chunk_lo + 1;
                    jr.need_r0 = max(jr.chunk_lo * kBlurChunkY - b, 0);
                    jr.need_r1 = min((jr.chunk_lo + jr.chunk_n) * kBlurChunkY + b, jr.bh);
                }
            } else {
                if (everywhere) { x0 = 0; y0 = t.band_y0; x1 = t.width; y1 = t.band_y0 + t.band_rows; }
                else {
                    x0 = max(jr.min_x, 0); x1 = min(jr.max_x + 1, t.width);
                    y0 = max(jr.min_y, t.band_y0); y1 = min(jr.max_y + 1, t.band_y0 + t.band_rows);
                }
                jr.cx0 = x0; jr.cy0 = y0; jr.cx1 = x1; jr.cy1 = y1;
            }
            if (x1 > x0 && y1 > y0) {
                jr.tx0 = x0 / kTile; jr.ty0 = y0 / kTile;
                jr.tw = (x1 - 1) / kTile - jr.tx0 + 1;
                jr.th = (y1 - 1) / kTile - jr.ty0 + 1;
                tiles = uint32_t(jr.tw) * uint32_t(jr.th);
            } else { jr.tx0 = jr.ty0 = 0; jr.tw = jr.th = 0; }
        }
        uint32_t total;
        uint32_t ex = block_exclusive_scan(tiles, sm, total);
        if (j < n) f.jobs[j].te_base = carry + ex;
        carry += total;
        if (j < n) {                                      // compact record for the tile compositor
            const job_rec &jr = f.jobs[j];
            const draw_rec &d = f.draws[jr.draw];
            comp_rec c;
            memset(&c, 0, sizeof c);
            c.kind = jr.kind; c.op = d.op;
            bool everywhere = jr.kind == JOB_CLIP || (jr.kind == JOB_MAIN && (~d.op & 8u));
            c.flags = (everywhere ? COMP_EVERYWHERE : 0u) | (jr.opaque ? COMP_OPAQUE : 0u);
            c.mask_src = d.mask_src; c.mask_dst = d.mask_dst; c.brush = d.brush; c.draw = jr.draw;
            c.te_base = jr.te_base;
            c.tx0 = jr.tx0; c.ty0 = jr.ty0; c.tw = jr.tw; c.th = jr.th;
            c.cx0 = jr.cx0; c.cy0 = jr.cy0; c.cx1 = jr.cx1; c.cy1 = jr.cy1;
            c.alpha = d.global_alpha;
            if (jr.kind == JOB_SHADOW) {
                for (int k = 0; k < 4; ++k) c.color[k] = d.shadow_color[k];
                c.border = jr.border; c.left = jr.left - jr.skew; c.top = jr.top; c.bw = jr.pitch;
            } else if (jr.kind == JOB_MAIN) {
                const brush_rec &b = f.brushes[d.brush];
                c.brush_type = b.n_colors ? b.type : 0xffu;     // 0xff: empty brush paints nothing
                if (b.type == CB200_BRUSH_COLOR && b.n_colors) {
                    float4 col = f.colors[b.first_color];
                    c.color[0] = col.x; c.color[1] = col.y; c.color[2] = col.z; c.color[3] = col.w;
                }
            }
            f.comp[j] = c;
            // compact hit-test record: tile box of the composite rectangle (empty box: tx1 < tx0)
            uint32_t bx0 = 1, by0 = 1, bx1 = 0, by1 = 0;
            if (jr.cx1 > jr.cx0 && jr.cy1 > jr.cy0) {
                bx0 = uint32_t(jr.cx0 / kTile); by0 = uint32_t(jr.cy0 / kTile);
                bx1 = uint32_t((jr.cx1 - 1) / kTile); by1 = uint32_t((jr.cy1 - 1) / kTile);
            }
            uint2 box;
            box.x = bx0 | by0 << 11 | (bx1 & 0x3ffu) << 22;
            box.y = bx1 >> 10 | by1 << 1 | uint32_t(jr.kind) << 12 | (everywhere ? JOBBOX_EVERYWHERE : 0u) |
                    (jr.opaque && jr.kind == JOB_MAIN ? JOBBOX_OPAQUE : 0u);
            f.job_box[j] = box;
            f.job_te[j] = jr.te_base;
        }
        // plane storage: order of allocation does not matter, only disjointness
        if (plane) {
            unsigned long long off = atomicAdd(&plane_carry, plane);
            atomicAdd(&working_carry, (unsigned long long)f.jobs[j].bw * (unsigned long long)(f.jobs[j].need_r1 - f.jobs[j].need_r0));
            f.jobs[j].plane_offset = off;
            f.comp[j].plane_lo = uint32_t(off); f.comp[j].plane_hi = uint32_t(off >> 32);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        h->n_tile_entries = carry;
        h->plane_floats = plane_carry;
        h->shadow_working_pixels = working_carry;
        if (carry > f.cap_tiles) atomicOr(&h->overflow, OVF_TILES);
        if (plane_carry > f.cap_planes) atomicOr(&h->overflow, OVF_PLANES);
    }
}

// zero the per-(tile entry, row) coverage bookkeeping
__global__ void k_clear_tiles(device_frame f)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    if (h->overflow) return;
    uint64_t n = uint64_t(h->n_tile_entries) * kTile;
    uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        f.te_backdrop[i] = 0.0f;
        f.te_first[i] = kNoRun;
        f.te_mask[i] = 0u;
        if (i < h->n_tile_entries) f.te_flags[i] = 0;
    }
}

}  // namespace

void launch_raster(const device_frame &f, const canvas_target &t, cudaStream_t s)
{
    launch_pdl(k_job_items, 1, kOneCta, 0, s, f);
    launch_pdl(k_edges, kGrid, kBlock, 0, s, f, t);
    launch_pdl(k_scan_rows, kGrid, kBlock, 0, s, f);
    launch_pdl(k_row_count, kGrid, kBlock, 0, s, f);
    launch_pdl(k_row_emit, kGrid, kBlock, 0, s, f);
    if (f.n_shadow_jobs) launch_pdl(k_shadow_boxes, kGrid, kBlock, 0, s, f, t);
    launch_pdl(k_job_tiles, 1, kOneCta, 0, s, f, t);
    launch_pdl(k_clear_tiles, kGrid, kBlock, 0, s, f);
}

}  // namespace cb200
