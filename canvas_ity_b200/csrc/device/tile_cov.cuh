// tile_cov.cuh -- coverage of one pixel of a 32-pixel tile row from what the row walk (coverage.cu) left
// per (tile entry, row):
//   backdrop  the running sum carried in from the left of the tile,
//   first     the index, in the per-scanline COMPACTED `cumulative` array, of the first pixel inside the tile that
//             holds a run (kNoRun: none; bit 31: the scanline's last run lies in this tile),
//   mask      bit p set: pixel p of the tile row holds a run.
// A lane (= pixel column) needs the sum after the last run at or left of its pixel: that is entry
// first + popc(mask & pixels up to mine) - 1, or the carried-in sum when there is none -- one popc and one load,
// no key compares, no shared memory, no warp synchronisation.  Result: the signed running sum; coverage is
// min(|sum|, 1) (reference render_main hpp:2570).
#pragma once

#include "frame.cuh"

namespace cb200 {

constexpr uint32_t kRowEnds = 0x80000000u;               // `first` bit 31, see coverage.cu

// kStopAfterLastRun: render_shadow walks the runs alone (hpp:2430-2452) -- after a scanline's LAST run only that
// run's own pixel takes the sum (to = x + 1), nothing carries on to the right.
template <bool kStopAfterLastRun>
__device__ __forceinline__ float pixel_sum(const float *cumulative, float backdrop, uint32_t first, uint32_t mask)
{
    if (first == kNoRun) return backdrop;                  // no edge touches this tile row
    const int lane = threadIdx.x & 31;
    if (kStopAfterLastRun && (first & kRowEnds) && lane > 31 - __clz(int(mask))) return 0.0f;
    const uint32_t upto = mask & (0xffffffffu >> (31 - lane));          // pixels <= mine that hold a run
    return upto ? cumulative[(first & ~kRowEnds) + uint32_t(__popc(upto)) - 1u] : backdrop;
}

}  // namespace cb200
