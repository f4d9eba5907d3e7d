// coverage.cu -- K6: per-scanline running coverage over the sorted runs and
// binning of that coverage into 32x32 tile entries.
//
// The reference coalesces equal (x, y) runs and keeps a running sum while it
// walks a scanline left to right (lines_to_runs hpp:2244-2252, render_main
// hpp:2570, 2601).  Here the thread that owns the first run of a (job, scanline)
// segment walks that segment in the same left-to-right order and
//   * stores the running sum after each pixel's last run, COMPACTED per scanline: entry
//     `head + k` of `cumulative` is the sum after the k-th distinct pixel of the segment that
//     starts at run `head` (runs of one pixel are coalesced exactly as the reference does),
//   * for every tile column the scanline enters, records the sum carried in from the left
//     (te_backdrop), the compacted index of the first pixel inside that tile (te_first; bit 31:
//     the scanline's last run lies in this tile) and a 32-bit mask of the tile's pixels that hold
//     a run (te_mask) -- so a lane of the tile compositor finds the coverage of its pixel with
//     one popc and one load: cumulative[first + popc(mask & pixels up to mine) - 1], or the
//     carried-in sum when no run lies at or left of it.  Nothing outside the tile is looked at.
// Segments are independent, so all scanlines of all jobs run in parallel; the
// serial part is one scanline's run list, as in the reference.
#include "frame.cuh"

namespace cb200 {

namespace {

constexpr uint32_t kShortRow = 24;      // longer scanline segments go to the warp-cooperative kernel (48 made the CTA wait at its barrier behind single long walks; 16 sent too many 17-24-run rows to a warp once k_rows_long prefetched: tiger coverage 0.127 -> 0.120 ms, batch of 2048 canvases 27.4 -> 26.9 ms)

// What is left over after the last run of a scanline.  render_main and clip (hpp:2551-2605, 3057-3099) walk the
// runs merged with the clip mask's, whose last run of a row sits at the right canvas edge: the residue carries on
// to that edge -- to the end of the job's tile rectangle here, and beyond it through a leak record when the
// rectangle (the outline's bounding box) stops short of the canvas.  render_shadow walks the runs alone
// (hpp:2430-2452): after a row's last run only that run's own pixel is painted, nothing carries on.
__device__ __forceinline__ void finish_row(const device_frame &f, frame_header *h, const job_rec &jr, uint32_t j, int y, int ly,
                                           bool binned, bool everywhere, uint32_t te_row, int c_prev, int c_end, float sum)
{
    if (!binned || sum == 0.0f || jr.kind == JOB_SHADOW || !(everywhere || fabsf(sum) >= kThreshold)) return;
    for (int cc = c_prev + 1; cc <= c_end; ++cc) {
        uint32_t te = te_row + uint32_t(cc - jr.tx0);
        f.te_backdrop[te * kTile + ly] = sum;
        { f.te_flags[te] = TE_NONEMPTY; f.te_job[te] = j; }
    }
    if (!everywhere && (c_end + 1) * kTile < h->width) {             // an everywhere job's rectangle is the whole canvas
        const uint32_t at = atomicAdd(&h->n_leaks, 1u);
        if (at < f.cap_leaks) {
            leak_rec l = { j, y, sum, 0u };
            f.leaks[at] = l;
            atomicOr(&f.job_box[j].y, JOBBOX_LEAKY);
        } else atomicOr(&h->overflow, OVF_LEAKS);
    }
}

constexpr uint32_t kRowEndsHere = 0x80000000u;          // te_first bit 31

// One short scanline segment, walked by one thread from its first run `i`.
__device__ __forceinline__ void walk_row(const device_frame &f, frame_header *h, const uint64_t *keys, const float *delta,
                                         uint32_t n, uint32_t i, uint32_t bx, uint32_t by)
{
    const uint64_t xmask = (1ull << bx) - 1, ymask = (1ull << by) - 1;
    uint64_t key = keys[i];
    const uint64_t row = key >> bx;
    uint32_t j = uint32_t(row >> by);
    if (j >= h->n_jobs) return;
    if (i + kShortRow < n && (keys[i + kShortRow] >> bx) == row) {   // sorted: the segment is longer than that
        f.long_rows[atomicAdd(&h->n_long_rows, 1u)] = i;             // a whole warp takes it
        return;
    }
    int y = int(row & ymask);
    const job_rec &jr = f.jobs[j];
    int ty = y / kTile - jr.ty0, ly = y % kTile;
    bool binned = ty >= 0 && ty < jr.th && jr.tw > 0;
    bool everywhere = jr.kind != JOB_MAIN || (~f.draws[jr.draw].op & 8u);
    uint32_t te_row = jr.te_base + uint32_t(ty) * uint32_t(jr.tw);
    int c_prev = jr.tx0 - 1;                                     // last tile column handled
    const int c_end = jr.tx0 + jr.tw - 1;
    float sum = 0.0f;
    uint32_t distinct = i;                                       // where the next distinct pixel's sum goes
    uint32_t open_slot = kNoRun, open_mask = 0, open_first = 0;  // the tile the walk is inside: its (te, row) slot
    // runs are fetched four at a time (plus the key that follows them), so that the walk waits for
    // memory once per batch instead of once per run; reading past the segment's end is harmless
    constexpr int kBatch = 4;
    bool more = true;
    for (uint32_t k0 = i; more; k0 += kBatch) {
        uint64_t kk[kBatch + 1];
        float dd[kBatch];
        kk[0] = key;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const uint32_t idx = k0 + uint32_t(u);
            kk[u + 1] = idx + 1 < n ? keys[idx + 1] : ~0ull;
            dd[u] = idx < n ? delta[idx] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (!more) break;
            int x = int(kk[u] & xmask);
            int c = x / kTile;
            if (binned && c > c_prev) {
                int last = min(c, c_end);
                if (open_slot != kNoRun) { f.te_first[open_slot] = open_first; f.te_mask[open_slot] = open_mask; open_slot = kNoRun; }
                // tiles entered since the previous run inherit the sum so far
                if (sum != 0.0f)
                    for (int cc = c_prev + 1; cc <= last; ++cc) {
                        uint32_t te = te_row + uint32_t(cc - jr.tx0);
                        f.te_backdrop[te * kTile + ly] = sum;
                        if (everywhere || fabsf(sum) >= kThreshold) { f.te_flags[te] = TE_NONEMPTY; f.te_job[te] = j; }
                    }
                if (c <= c_end) {
                    uint32_t te = te_row + uint32_t(c - jr.tx0);
                    open_slot = te * kTile + uint32_t(ly); open_first = distinct; open_mask = 0;
                    { f.te_flags[te] = TE_NONEMPTY; f.te_job[te] = j; }
                }
                c_prev = max(c_prev, last);
            }
            sum += dd[u];
            const uint64_t nkey = kk[u + 1];
            const bool same_row = (nkey >> bx) == row;
            if (!(same_row && int(nkey & xmask) == x)) {             // the pixel's last run: its sum is final
                f.cumulative[distinct++] = sum;
                open_mask |= 1u << (x % kTile);
            }
            more = same_row;
        }
        key = kk[kBatch];
    }
    if (open_slot != kNoRun) { f.te_first[open_slot] = open_first | kRowEndsHere; f.te_mask[open_slot] = open_mask; }
    finish_row(f, h, jr, j, y, ly, binned, everywhere, te_row, c_prev, c_end, sum);
}

// A CTA takes 1024 consecutive runs per step, finds the segment heads among them (coalesced key
// compares), compacts their indices in shared memory and then hands one head to each thread: all
// lanes of a warp walk a segment, instead of the one lane in four that happens to sit on a head
// (4x fewer warp instructions on the tiger; what counts when several frames share the GPU).
constexpr int kRowKeys = 4;
constexpr int kRowStep = kBlock * kRowKeys;

__global__ void __launch_bounds__(kBlock) k_rows(device_frame f, int sb)
{
    grid_dependency_wait();
    __shared__ uint32_t heads[kRowStep];
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t n = h->n_runs;
    const uint32_t bx = h->sort_bits_x, by = h->sort_bits_y;
    const uint64_t *keys = f.keys[sb];
    const float *delta = f.vals[sb];
    for (uint32_t step = blockIdx.x * kRowStep; step < n; step += gridDim.x * kRowStep) {
        uint32_t is_head = 0;                                     // bit k: my k-th run of this step starts a segment
#pragma unroll
        for (int k = 0; k < kRowKeys; ++k) {
            const uint32_t i = step + uint32_t(k) * kBlock + threadIdx.x;
            if (i < n && (i == 0 || (keys[i - 1] >> bx) != (keys[i] >> bx))) is_head |= 1u << k;
        }
        uint32_t total;
        uint32_t at = block_exclusive_scan(uint32_t(__popc(is_head)), sm, total);
#pragma unroll
        for (int k = 0; k < kRowKeys; ++k)
            if (is_head >> k & 1u) heads[at++] = step + uint32_t(k) * kBlock + threadIdx.x;
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < total; q += kBlock) walk_row(f, h, keys, delta, n, heads[q], bx, by);
        __syncthreads();
    }
}

// Long scanline segments (thin, nearly horizontal shapes put thousands of runs on
// one scanline): one warp per segment, 32 runs per step, running sum by a warp
// prefix scan carried in double precision (so the float result does not depend on
// the scan's association order), same bookkeeping as k_rows.
__global__ void __launch_bounds__(kBlock) k_rows_long(device_frame f, int sb)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t n = h->n_runs, n_long = h->n_long_rows;
    const uint32_t bx = h->sort_bits_x, by = h->sort_bits_y;
    const uint64_t xmask = (1ull << bx) - 1, ymask = (1ull << by) - 1;
    const uint64_t *keys = f.keys[sb];
    const float *delta = f.vals[sb];
    const float quiet_nan = __int_as_float(0x7fc00000);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = warp; w < n_long; w += n_warps) {
        const uint32_t head = f.long_rows[w];
        const uint64_t row = keys[head] >> bx;
        const uint32_t j = uint32_t(row >> by);
        const int y = int(row & ymask);
        const job_rec &jr = f.jobs[j];
        const int ty = y / kTile - jr.ty0, ly = y % kTile;
        const bool binned = ty >= 0 && ty < jr.th && jr.tw > 0;
        const bool everywhere = jr.kind != JOB_MAIN || (~f.draws[jr.draw].op & 8u);
        const uint32_t te_row = jr.te_base + uint32_t(ty) * uint32_t(jr.tw);
        const int c_end = jr.tx0 + jr.tw - 1;
        int c_carry = jr.tx0 - 1;                     // tile column of the last run seen so far
        double carry = 0.0;
        uint32_t distinct = head;                     // compacted index of the next distinct pixel (see walk_row)
        uint32_t last_slot = kNoRun;                  // (te, row) slot of the tile that holds the latest run
        // the next 32 runs are requested while these are processed (the walk waited for memory once per step: a
        // third of the kernel's stall samples on a batch, where nearly every scanline is a long one)
        uint64_t key_ahead = head + uint32_t(lane) < n ? keys[head + uint32_t(lane)] : ~0ull;
        float delta_ahead = head + uint32_t(lane) < n ? delta[head + uint32_t(lane)] : 0.0f;
        for (uint32_t base = head;; base += 32) {
            uint32_t idx = base + uint32_t(lane);
            const uint64_t key = key_ahead;
            const float dv = delta_ahead;
            key_ahead = idx + 32 < n ? keys[idx + 32] : ~0ull;
            delta_ahead = idx + 32 < n ? delta[idx + 32] : 0.0f;
            bool valid = (key >> bx) == row;
            // the key that follows mine: my neighbour's, or the first of the next step
            const uint64_t key_right = __shfl_down_sync(0xffffffffu, key, 1), key_next_step = __shfl_sync(0xffffffffu, key_ahead, 0);
            const uint64_t nkey = lane == 31 ? key_next_step : key_right;
            double v = valid ? double(dv) : 0.0;
            double incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            float after = float(carry + incl), before = float(carry + incl - v);
            int x = int(key & xmask), c = valid ? x / kTile : 0x3fffffff;
            int c_left = __shfl_up_sync(0xffffffffu, c, 1);
            if (lane == 0) c_left = c_carry;
            // distinct pixels: a run is its pixel's last when the next run is elsewhere
            const bool final_of_pixel = valid && !((nkey >> bx) == row && int(nkey & xmask) == x);
            const uint32_t finals = __ballot_sync(0xffffffffu, final_of_pixel);
            const uint32_t my_entry = distinct + uint32_t(__popc(finals & ((1u << lane) - 1u)));   // entry of my pixel
            if (final_of_pixel) f.cumulative[my_entry] = after;
            const bool in_rect = valid && binned && c >= jr.tx0 && c <= c_end;    // projected runs may lie left of a shadow's rectangle
            const uint32_t slot = in_rect ? (te_row + uint32_t(c - jr.tx0)) * kTile + uint32_t(ly) : kNoRun;
            if (valid) {
                if (binned && c > c_left) {
                    int last = min(c, c_end);
                    if (before != 0.0f)
                        for (int cc = c_left + 1; cc <= last; ++cc) {
                            uint32_t te = te_row + uint32_t(cc - jr.tx0);
                            f.te_backdrop[te * kTile + ly] = before;
                            if (everywhere || fabsf(before) >= kThreshold) { f.te_flags[te] = TE_NONEMPTY; f.te_job[te] = j; }
                        }
                    if (c <= c_end) {
                        uint32_t te = te_row + uint32_t(c - jr.tx0);
                        f.te_first[slot] = my_entry;                  // first run inside the tile: its pixel's entry
                        { f.te_flags[te] = TE_NONEMPTY; f.te_job[te] = j; }
                    }
                }
            }
            // pixel masks: the lanes of one tile combine their bits, one of them adds them to the slot (zeroed by
            // k_clear_tiles; this warp is the only writer of the row)
            const uint32_t peers = __match_any_sync(0xffffffffu, slot);
            const uint32_t bits = __reduce_or_sync(peers, final_of_pixel ? 1u << (x % kTile) : 0u);
            if (slot != kNoRun && lane == __ffs(int(peers)) - 1 && bits) atomicOr(&f.te_mask[slot], bits);   // fire and forget (a load-or-store waited for the load)
            uint32_t vm = __ballot_sync(0xffffffffu, valid);
            int last_lane = 31 - __clz(int(vm));
            carry += __shfl_sync(0xffffffffu, incl, last_lane);
            c_carry = max(c_carry, min(__shfl_sync(0xffffffffu, c, last_lane), c_end));
            distinct += uint32_t(__popc(finals));
            const uint32_t tail_slot = __shfl_sync(0xffffffffu, slot, last_lane);
            if (vm) last_slot = tail_slot;
            __syncwarp();                                 // slot updates of this step are visible to the next step's lanes
            if (vm != 0xffffffffu) break;
        }
        if (lane == 0 && last_slot != kNoRun) f.te_first[last_slot] |= kRowEndsHere;
        if (lane == 0) finish_row(f, h, jr, j, y, ly, binned, everywhere, te_row, c_carry, c_end, float(carry));
    }
}

// A tile entry is "covered" when every scanline of the tile that lies on the canvas
// has full coverage carried in from the left and no run inside the tile: the draw
// paints all of it with coverage exactly 1.  The compositor uses this for occlusion
// culling (an opaque covered draw makes everything beneath it irrelevant).
__global__ void __launch_bounds__(kBlock) k_tile_flags(device_frame f, canvas_target t)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = h->n_tile_entries;
    // one warp per tile entry, one lane per scanline; te_job remembers which job marked it
    for (uint32_t te = warp; te < n; te += n_warps) {
        uint32_t flags = f.te_flags[te];
        if (!(flags & TE_NONEMPTY)) continue;
        uint32_t lo = f.te_job[te];
        const job_rec &jr = f.jobs[lo];
        if (jr.kind != JOB_MAIN || !jr.opaque) continue;
        uint32_t local = te - jr.te_base;
        int ty = jr.ty0 + int(local / uint32_t(jr.tw));
        int y = ty * kTile + lane;
        bool on_canvas = y >= t.band_y0 && y < t.band_y0 + t.band_rows;
        bool full = !on_canvas || (f.te_first[te * kTile + lane] == kNoRun && fabsf(f.te_backdrop[te * kTile + lane]) >= 1.0f);
        if (__all_sync(0xffffffffu, full) && lane == 0) {
            f.te_flags[te] = flags | TE_COVERED;
            // the compositor's occlusion culling: the LAST covering job of the target tile
            const int tx = jr.tx0 + int(local % uint32_t(jr.tw));
            const int tiles_x = (t.width + kTile - 1) / kTile;
            const int row = t.n_canvases > 1 ? int(jr.canvas) * (t.slot_rows / kTile) + ty : ty - t.band_y0 / kTile;
            atomicMax(&f.tile_cover[size_t(row) * size_t(tiles_x) + size_t(tx)], lo + 1u);
        }
    }
}

}  // namespace

void launch_rows(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    launch_pdl(k_rows, kGrid, kBlock, 0, s, f, sorted_buffer);
    launch_pdl(k_rows_long, kGrid, kBlock, 0, s, f, sorted_buffer);
    launch_pdl(k_tile_flags, kGrid, kBlock, 0, s, f, t);
}

}  // namespace cb200
