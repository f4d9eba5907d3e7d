// tile_cov.cuh -- dense coverage of one 32-pixel tile row from the sorted runs.
//
// One warp owns one scanline of a tile, one lane per pixel.  The row-walk kernel
// (coverage.cu) left, per (tile entry, row): the running sum carried in from the
// left (`backdrop`) and the index of the first run inside the tile.  The warp
// scatters the tile's running sums into a 32-float shared-memory row and
// forward-fills it with one ballot + one shuffle -- the "per-scanline coverage
// accumulation using warp-shuffle prefix scans and shared-memory tile staging" of
// the north star.  Result: signed running sum for this lane's pixel; coverage is
// min(|sum|, 1) (reference render_main hpp:2570).
#pragma once

#include "frame.cuh"

namespace cb200 {

struct cov_source {
    const uint64_t *keys;
    const float *cumulative;
    const float *backdrop;
    const uint32_t *first;
    uint32_t n_runs, bx, by;
};

__device__ __forceinline__ cov_source make_cov_source(const device_frame &f, int sb)
{
    cov_source c;
    c.keys = f.keys[sb];
    c.cumulative = f.cumulative;
    c.backdrop = f.te_backdrop;
    c.first = f.te_first;
    c.n_runs = f.hdr->n_runs;
    c.bx = f.hdr->sort_bits_x;
    c.by = f.hdr->sort_bits_y;
    return c;
}

// `row_buf`: 32 floats of shared memory private to the calling warp.
// (job, y) identify the scanline, x0 is the tile's left pixel in raster space.
template <bool kStopAfterLastRun>
__device__ __forceinline__ float tile_row_sum(const cov_source &c, uint32_t te, int ly, uint32_t job,
                                              int y, int x0, float *row_buf)
{
    const int lane = threadIdx.x & 31;
    const float backdrop = c.backdrop[te * kTile + ly];
    const uint32_t first = c.first[te * kTile + ly];
    if (first == kNoRun) return backdrop;                   // no edge touches this tile row
    const uint64_t row_key = (uint64_t(job) << c.by) | uint64_t(uint32_t(y));
    const uint64_t xmask = (1ull << c.bx) - 1;
    const float quiet_nan = __int_as_float(0x7fc00000);
    row_buf[lane] = quiet_nan;
    __syncwarp();
    bool row_goes_on = false;                               // runs of this row to the right of the tile
    for (uint32_t k = first;; k += 32) {
        uint32_t idx = k + uint32_t(lane);
        bool ok = idx < c.n_runs;
        uint64_t key = ok ? c.keys[idx] : ~0ull;
        int col = int(key & xmask) - x0;
        const bool same_row = ok && (key >> c.bx) == row_key;
        ok = same_row && col < kTile;
        if (ok) {
            float v = c.cumulative[idx];
            if (v == v) row_buf[col] = v;                   // NaN = superseded by a later run
        }
        const uint32_t stop = ~__ballot_sync(0xffffffffu, ok);
        if (stop) { row_goes_on = __shfl_sync(0xffffffffu, same_row, __ffs(int(stop)) - 1); break; }
    }
    __syncwarp();
    float mine = row_buf[lane];
    uint32_t have = __ballot_sync(0xffffffffu, mine == mine);
    uint32_t upto = have & (0xffffffffu >> (31 - lane));    // pixels <= mine that hold a sum
    int src = upto ? 31 - __clz(upto) : 0;
    float got = __shfl_sync(0xffffffffu, mine, src);
    __syncwarp();
    // render_shadow walks the runs alone (hpp:2430-2452): after the row's LAST run only that run's own pixel takes
    // the sum (to = x + 1), nothing carries on to the right
    if (kStopAfterLastRun && !row_goes_on && have && lane > 31 - __clz(have)) return 0.0f;
    return upto ? got : backdrop;
}

}  // namespace cb200
