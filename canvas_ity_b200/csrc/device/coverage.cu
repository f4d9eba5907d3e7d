// coverage.cu -- K6: per-scanline running coverage over the sorted runs and
// binning of that coverage into 32x32 tile entries.
//
// The reference coalesces equal (x, y) runs and keeps a running sum while it
// walks a scanline left to right (lines_to_runs hpp:2244-2252, render_main
// hpp:2570, 2601).  Here the thread that owns the first run of a (job, scanline)
// segment walks that segment in the same left-to-right order and
//   * stores the running sum after each pixel's last run (`cumulative`; earlier
//     runs of the same pixel get NaN = "superseded"),
//   * for every tile column the scanline enters, records the sum carried in from
//     the left (te_backdrop) and the index of the first run inside that tile
//     (te_first), so the tile compositor can rebuild dense coverage for its 32
//     pixels without looking at anything outside the tile.
// Segments are independent, so all scanlines of all jobs run in parallel; the
// serial part is one scanline's run list, as in the reference.
#include "frame.cuh"

namespace cb200 {

namespace {

__global__ void __launch_bounds__(kBlock) k_rows(device_frame f, int sb)
{
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t n = h->n_runs;
    const uint32_t bx = h->sort_bits_x, by = h->sort_bits_y;
    const uint64_t xmask = (1ull << bx) - 1, ymask = (1ull << by) - 1;
    const uint64_t *keys = f.keys[sb];
    const float *delta = f.vals[sb];
    const float quiet_nan = __int_as_float(0x7fc00000);
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t row = keys[i] >> bx;
        if (i > 0 && (keys[i - 1] >> bx) == row) continue;          // not a segment head
        uint32_t j = uint32_t(row >> by);
        int y = int(row & ymask);
        const job_rec &jr = f.jobs[j];
        int ty = y / kTile - jr.ty0, ly = y % kTile;
        bool binned = ty >= 0 && ty < jr.th && jr.tw > 0;
        bool everywhere = jr.kind != JOB_MAIN || (~f.draws[jr.draw].op & 8u);
        uint32_t te_row = jr.te_base + uint32_t(ty) * uint32_t(jr.tw);
        int c_prev = jr.tx0 - 1;                                     // last tile column handled
        const int c_end = jr.tx0 + jr.tw - 1;
        float sum = 0.0f;
        uint32_t k = i;
        uint64_t key = keys[k];
        for (;;) {
            int x = int(key & xmask);
            int c = x / kTile;
            if (binned && c > c_prev) {
                int last = min(c, c_end);
                // tiles entered since the previous run inherit the sum so far
                if (sum != 0.0f)
                    for (int cc = c_prev + 1; cc <= last; ++cc) {
                        uint32_t te = te_row + uint32_t(cc - jr.tx0);
                        f.te_backdrop[te * kTile + ly] = sum;
                        if (everywhere || fabsf(sum) >= kThreshold) f.te_flags[te] = 1;
                    }
                if (c <= c_end) {
                    uint32_t te = te_row + uint32_t(c - jr.tx0);
                    f.te_first[te * kTile + ly] = k;
                    f.te_flags[te] = 1;
                }
                c_prev = max(c_prev, last);
            }
            sum += delta[k];
            uint32_t nk = k + 1;
            uint64_t nkey = nk < n ? keys[nk] : ~0ull;
            bool same_row = (nkey >> bx) == row;
            f.cumulative[k] = (same_row && int(nkey & xmask) == x) ? quiet_nan : sum;
            if (!same_row) break;
            k = nk;
            key = nkey;
        }
        // whatever is left over after the last run spills to the right edge
        if (binned && sum != 0.0f && (everywhere || fabsf(sum) >= kThreshold))
            for (int cc = c_prev + 1; cc <= c_end; ++cc) {
                uint32_t te = te_row + uint32_t(cc - jr.tx0);
                f.te_backdrop[te * kTile + ly] = sum;
                f.te_flags[te] = 1;
            }
    }
}

// A tile entry is "covered" when every scanline of the tile that lies on the canvas
// has full coverage carried in from the left and no run inside the tile: the draw
// paints all of it with coverage exactly 1.  The compositor uses this for occlusion
// culling (an opaque covered draw makes everything beneath it irrelevant).
__global__ void __launch_bounds__(kBlock) k_tile_flags(device_frame f, canvas_target t)
{
    frame_header *h = f.hdr;
    if (h->overflow) return;
    const uint32_t n_jobs = h->n_jobs;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = h->n_tile_entries;
    // one warp per tile entry, one lane per scanline; entry -> job by binary search on te_base
    for (uint32_t te = warp; te < n; te += n_warps) {
        uint32_t flags = f.te_flags[te];
        if (!(flags & TE_NONEMPTY)) continue;
        uint32_t lo = 0, hi = n_jobs;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (f.jobs[mid].te_base <= te) lo = mid; else hi = mid;
        }
        while (lo + 1 < n_jobs && f.jobs[lo].tw * f.jobs[lo].th == 0) ++lo;      // skip empty jobs sharing the base
        const job_rec &jr = f.jobs[lo];
        if (jr.kind != JOB_MAIN || !jr.opaque) continue;
        uint32_t local = te - jr.te_base;
        int ty = jr.ty0 + int(local / uint32_t(jr.tw));
        int y = ty * kTile + lane;
        bool on_canvas = y >= t.band_y0 && y < t.band_y0 + t.band_rows;
        bool full = !on_canvas || (f.te_first[te * kTile + lane] == kNoRun && fabsf(f.te_backdrop[te * kTile + lane]) >= 1.0f);
        if (__all_sync(0xffffffffu, full) && lane == 0) f.te_flags[te] = flags | TE_COVERED;
    }
}

}  // namespace

void launch_rows(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s)
{
    k_rows<<<kGrid, kBlock, 0, s>>>(f, sorted_buffer);
    k_tile_flags<<<kGrid, kBlock, 0, s>>>(f, t);
}

}  // namespace cb200
