// sort.cu -- K5: stable LSD radix sort of the pixel runs by key = job|y|x
// (replaces the std::sort of reference lines_to_runs, hpp:2243; the tie order
// among runs on the same pixel is irrelevant to the summed coverage).
//
// Digits are up to 9 bits wide: ceil(key_bits / 9) passes of equal width; the host knows
// key_bits from the canvas size and the job count (34 bits = 4 passes for the 4096x4096 tiger).  Per pass, three launches with fixed grids
// (the run count lives on the device):
//   k_sort_hist     per-CTA digit histogram of its slice (digit-major table)
//   k_sort_scan     one CTA per digit: exclusive scan across the CTAs + digit total
//   k_sort_scatter  each CTA re-reads its slice in order, 2048 keys (8 per thread) per step;
//                   ranks are made stable with warp match + per-warp digit counts; the step is
//                   regrouped by digit in shared memory before it is stored
// Frames of many small jobs (batches of small canvases) are sorted job by job instead, each job's range inside one
// CTA and by its (y, x) bits only: k_job_runs + k_sort_jobs (see segmented_sort_applies below).
#include "frame.cuh"

namespace cb200 {

namespace {

constexpr int kMaxRadix = 512;          // digits are up to 9 bits wide

// ---- which path sorts the frame ---------------------------------------------------------------------------------
// Runs are EMITTED job by job (K4's row items are numbered by piece slot, piece slots by job), so the job field
// of the unsorted keys is already non-decreasing: a frame of many small jobs -- a batch of small canvases -- only
// needs every job's own range sorted by (y, x).  k_sort_jobs does that CTA-locally (below); it applies when the host
// enabled it for the frame (job_run_begin set: many jobs, pass parity matches) and the largest job is at most twice a
// CTA's fair share of the runs, so that the work balances.  Every kernel of both paths evaluates the same rule.
__device__ __forceinline__ bool segmented_sort_applies(const device_frame &f)
{
    const frame_header *h = f.hdr;
    return f.job_run_begin != nullptr && h->max_job_runs != 0 &&
           uint64_t(h->max_job_runs) * uint64_t(kGrid / 2) <= uint64_t(h->n_runs);
}

__device__ __forceinline__ void sort_slice(uint32_t n, uint32_t &begin, uint32_t &end)
{
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kBlock - 1) / kBlock * kBlock;
    uint64_t b = uint64_t(per) * blockIdx.x;
    begin = b < n ? uint32_t(b) : n;
    end = b + per < n ? uint32_t(b + per) : n;
}

__global__ void __launch_bounds__(kBlock) k_sort_hist(device_frame f, int src, int shift, int width)
{
    grid_dependency_wait();
    __shared__ uint32_t bins[kMaxRadix];
    frame_header *h = f.hdr;
    if (segmented_sort_applies(f)) return;
    uint32_t n = h->overflow ? 0 : h->n_runs, begin, end;
    sort_slice(n, begin, end);
    const uint32_t radix = 1u << width, mask = radix - 1;
    for (uint32_t d = threadIdx.x; d < radix; d += kBlock) bins[d] = 0;
    __syncthreads();
    const uint64_t *keys = f.keys[src];
    for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock)
        atomicAdd(&bins[uint32_t(keys[i] >> shift) & mask], 1u);
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < radix; d += kBlock) f.sort_hist[d * kGrid + blockIdx.x] = bins[d];   // digit-major
}

// One CTA per digit: exclusive scan of that digit's per-CTA counts (coalesced),
// digit total to sort_hist[kMaxRadix * kGrid + digit].
__global__ void __launch_bounds__(kBlock) k_sort_scan(device_frame f)
{
    grid_dependency_wait();
    __shared__ uint32_t sm[33];
    if (segmented_sort_applies(f)) return;
    uint32_t *row = f.sort_hist + blockIdx.x * kGrid;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < kGrid; base += kBlock) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < kGrid ? row[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, sm, total);
        if (i < kGrid) row[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) f.sort_hist[kMaxRadix * kGrid + blockIdx.x] = carry;
}

// Keys per thread and step of the scatter: a CTA ranks 2048 keys between two barriers.  Warp w
// owns keys [w * 256, (w + 1) * 256) of the step, lane l its keys k * 32 + l, so "warp, then k,
// then lane" is input order and the ranks below are stable.
constexpr int kSortKeys = 8;
constexpr int kSortStep = kBlock * kSortKeys;

__global__ void __launch_bounds__(kBlock) k_sort_scatter(device_frame f, int src, int shift, int width)
{
    grid_dependency_wait();
    __shared__ uint32_t base[kMaxRadix];             // next free global slot per digit for this CTA
    __shared__ uint32_t step_base[kMaxRadix];        // the same at the start of the current step
    __shared__ uint32_t local_start[kMaxRadix];      // first slot of the digit in the staged step
    __shared__ uint32_t warp_count[kBlock / 32][kMaxRadix];
    __shared__ uint64_t staged_key[kSortStep];       // the step's keys grouped by digit, input order within
    __shared__ float staged_val[kSortStep];
    __shared__ uint32_t sm[33];
    frame_header *h = f.hdr;
    if (segmented_sort_applies(f)) return;
    uint32_t n = h->overflow ? 0 : h->n_runs, begin, end;
    sort_slice(n, begin, end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t radix = 1u << width, mask = radix - 1;
    const uint32_t d0 = 2 * threadIdx.x, d1 = d0 + 1;        // thread t owns digits 2t and 2t+1
    {   // global base of digit d = keys with a smaller digit + this digit's keys in earlier CTAs
        uint32_t t0 = d0 < radix ? f.sort_hist[kMaxRadix * kGrid + d0] : 0;
        uint32_t t1 = d1 < radix ? f.sort_hist[kMaxRadix * kGrid + d1] : 0;
        uint32_t total;
        uint32_t below = block_exclusive_scan(t0 + t1, sm, total);
        if (d0 < radix) base[d0] = below + f.sort_hist[d0 * kGrid + blockIdx.x];
        if (d1 < radix) base[d1] = below + t0 + f.sort_hist[d1 * kGrid + blockIdx.x];
    }
    for (uint32_t d = threadIdx.x; d < radix; d += kBlock)
        for (int w = 0; w < kBlock / 32; ++w) warp_count[w][d] = 0;
    __syncthreads();
    const uint64_t *kin = f.keys[src];
    const float *vin = f.vals[src];
    uint64_t *kout = f.keys[src ^ 1];
    float *vout = f.vals[src ^ 1];
    uint32_t *mine = warp_count[warp];
    for (uint32_t step = begin; step < end; step += kSortStep) {
        uint64_t key[kSortKeys];
        float val[kSortKeys];
        uint32_t rank[kSortKeys];                    // rank among the warp's keys of the same digit
        const uint32_t first = step + uint32_t(warp) * (32 * kSortKeys) + uint32_t(lane);
        const uint32_t in_step = min(end - step, uint32_t(kSortStep));
#pragma unroll
        for (int k = 0; k < kSortKeys; ++k) {
            const uint32_t i = first + uint32_t(k) * 32;
            const bool valid = i < end;
            key[k] = valid ? kin[i] : ~0ull;
            val[k] = valid ? vin[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < kSortKeys; ++k) {
            const bool valid = first + uint32_t(k) * 32 < end;
            const uint32_t digit = uint32_t(key[k] >> shift) & mask;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : kMaxRadix + uint32_t(lane));
            const uint32_t ahead = __popc(peers & ((1u << lane) - 1u));
            const uint32_t seen = valid ? mine[digit] : 0;
            rank[k] = seen + ahead;
            __syncwarp();
            if (valid && ahead == 0) mine[digit] = seen + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // each digit: prefix over the warps, advance the CTA base, count the step's keys
        uint32_t c0 = 0, c1 = 0;
        for (uint32_t d = d0; d <= d1 && d < radix; ++d) {
            const uint32_t was = base[d];
            uint32_t run = was;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) {
                uint32_t c = warp_count[w][d];
                warp_count[w][d] = run;
                run += c;
            }
            base[d] = run;
            step_base[d] = was;
            (d == d0 ? c0 : c1) = run - was;
        }
        uint32_t total;
        const uint32_t before = block_exclusive_scan(c0 + c1, sm, total);     // barriers inside
        if (d0 < radix) local_start[d0] = before;
        if (d1 < radix) local_start[d1] = before + c0;
        __syncthreads();
        // stage the step grouped by digit ...
#pragma unroll
        for (int k = 0; k < kSortKeys; ++k) {
            if (first + uint32_t(k) * 32 < end) {
                const uint32_t digit = uint32_t(key[k] >> shift) & mask;
                const uint32_t slot = local_start[digit] + (mine[digit] + rank[k] - step_base[digit]);
                staged_key[slot] = key[k];
                staged_val[slot] = val[k];
            }
        }
        __syncthreads();
        // ... so that neighbouring threads store neighbouring keys of one digit: whole sectors
        // instead of one 8-byte fragment per key
        for (uint32_t i = threadIdx.x; i < in_step; i += kBlock) {
            const uint64_t kk = staged_key[i];
            const uint32_t digit = uint32_t(kk >> shift) & mask;
            const uint32_t dst = step_base[digit] + (i - local_start[digit]);
            kout[dst] = kk;
            vout[dst] = staged_val[i];
        }
        __syncthreads();
        // prefixes (also of digits absent from a warp) must not leak into the next step
        for (uint32_t d = threadIdx.x; d < radix; d += kBlock) {
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) warp_count[w][d] = 0;
        }
        __syncthreads();
    }
}

// ---- job-segmented sort ------------------------------------------------------------------------------------------

// First run of every job in the unsorted keys (lower bound of the job field, which is non-decreasing there), and the
// size of the largest job.  One thread per job.
__global__ void __launch_bounds__(kBlock) k_job_runs(device_frame f)
{
    grid_dependency_wait();
    frame_header *h = f.hdr;
    const uint32_t n_jobs = h->n_jobs, n = h->overflow ? 0 : h->n_runs;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t shift = h->sort_bits_x + h->sort_bits_y;
    const uint64_t *keys = f.keys[0];
    auto first_of = [&](uint32_t job) -> uint32_t {
        uint32_t lo = 0, hi = n;                                  // first index whose job field is >= job
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (uint32_t(keys[mid] >> shift) < job) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    uint32_t size = 0;
    if (j < n_jobs) {
        const uint32_t a = first_of(j), b = first_of(j + 1);
        f.job_run_begin[j] = a;
        if (j == n_jobs - 1) f.job_run_begin[n_jobs] = b;
        size = b - a;
    }
    size = __reduce_max_sync(0xffffffffu, size);
    if ((threadIdx.x & 31) == 0 && size) atomicMax(&h->max_job_runs, size);
}

// Every job's range sorted by (y, x): stable LSD radix passes like the global ones -- same digit ranking, same staging
// of a step by digit -- but histogram, scan and scatter of a range all happen inside the CTA that took the job off the
// ticket, so the passes need no grid-wide hand-over, ping-pong between the two key buffers inside the job's own range
// (a few hundred KB: L2), and only cover the (y, x) bits: 2 passes for a 256^2 canvas where the global sort needs 4.
// The order it leaves is exactly the global stable sort's order (job-major on input, stable passes), bit for bit.
__global__ void __launch_bounds__(kBlock, 4) k_sort_jobs(device_frame f, int passes, int width)
{
    grid_dependency_wait();
    __shared__ uint32_t base[kMaxRadix];             // next free global slot per digit
    __shared__ uint32_t step_base[kMaxRadix];        // the same at the start of the current step
    __shared__ uint32_t local_start[kMaxRadix];      // first slot of the digit in the staged step
    __shared__ uint32_t warp_count[kBlock / 32][kMaxRadix];
    __shared__ uint64_t staged_key[kSortStep];
    __shared__ float staged_val[kSortStep];
    __shared__ uint32_t sm[33];
    __shared__ uint32_t taken;
    frame_header *h = f.hdr;
    if (h->overflow || !segmented_sort_applies(f)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t radix = 1u << width, mask = radix - 1;
    const uint32_t d0 = 2 * threadIdx.x, d1 = d0 + 1;        // thread t owns digits 2t and 2t+1
    const uint32_t n_jobs = h->n_jobs;
    uint32_t *mine = warp_count[warp];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) taken = atomicAdd(&h->tickets[8], 1u);
        __syncthreads();
        const uint32_t job = taken;
        if (job >= n_jobs) break;
        const uint32_t begin = f.job_run_begin[job], end = f.job_run_begin[job + 1];
        if (end <= begin) continue;
        for (int p = 0; p < passes; ++p) {
            const int src = p & 1, shift = p * width;
            const uint64_t *kin = f.keys[src];
            const float *vin = f.vals[src];
            uint64_t *kout = f.keys[src ^ 1];
            float *vout = f.vals[src ^ 1];
            // digit histogram of the range (step_base doubles as the bins)
            for (uint32_t d = threadIdx.x; d < radix; d += kBlock) {
                step_base[d] = 0;
                for (int w = 0; w < kBlock / 32; ++w) warp_count[w][d] = 0;
            }
            __syncthreads();
            for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock)
                atomicAdd(&step_base[uint32_t(kin[i] >> shift) & mask], 1u);
            __syncthreads();
            {   // base of digit d = range start + keys of the range with a smaller digit
                const uint32_t t0 = d0 < radix ? step_base[d0] : 0, t1 = d1 < radix ? step_base[d1] : 0;
                uint32_t total;
                const uint32_t below = block_exclusive_scan(t0 + t1, sm, total);
                if (d0 < radix) base[d0] = begin + below;
                if (d1 < radix) base[d1] = begin + below + t0;
            }
            __syncthreads();
            for (uint32_t step = begin; step < end; step += kSortStep) {
                uint64_t key[kSortKeys];
                float val[kSortKeys];
                uint32_t rank[kSortKeys];                    // rank among the warp's keys of the same digit
                const uint32_t first = step + uint32_t(warp) * (32 * kSortKeys) + uint32_t(lane);
                const uint32_t in_step = min(end - step, uint32_t(kSortStep));
#pragma unroll
                for (int k = 0; k < kSortKeys; ++k) {
                    const uint32_t i = first + uint32_t(k) * 32;
                    const bool valid = i < end;
                    key[k] = valid ? kin[i] : ~0ull;
                    val[k] = valid ? vin[i] : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < kSortKeys; ++k) {
                    const bool valid = first + uint32_t(k) * 32 < end;
                    const uint32_t digit = uint32_t(key[k] >> shift) & mask;
                    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : kMaxRadix + uint32_t(lane));
                    const uint32_t ahead = __popc(peers & ((1u << lane) - 1u));
                    const uint32_t seen = valid ? mine[digit] : 0;
                    rank[k] = seen + ahead;
                    __syncwarp();
                    if (valid && ahead == 0) mine[digit] = seen + __popc(peers);
                    __syncwarp();
                }
                __syncthreads();
                // each digit: prefix over the warps, advance the base, count the step's keys
                uint32_t c0 = 0, c1 = 0;
                for (uint32_t d = d0; d <= d1 && d < radix; ++d) {
                    const uint32_t was = base[d];
                    uint32_t run = was;
#pragma unroll
                    for (int w = 0; w < kBlock / 32; ++w) {
                        uint32_t c = warp_count[w][d];
                        warp_count[w][d] = run;
                        run += c;
                    }
                    base[d] = run;
                    step_base[d] = was;
                    (d == d0 ? c0 : c1) = run - was;
                }
                uint32_t total;
                const uint32_t before = block_exclusive_scan(c0 + c1, sm, total);     // barriers inside
                if (d0 < radix) local_start[d0] = before;
                if (d1 < radix) local_start[d1] = before + c0;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < kSortKeys; ++k) {
                    if (first + uint32_t(k) * 32 < end) {
                        const uint32_t digit = uint32_t(key[k] >> shift) & mask;
                        const uint32_t slot = local_start[digit] + (mine[digit] + rank[k] - step_base[digit]);
                        staged_key[slot] = key[k];
                        staged_val[slot] = val[k];
                    }
                }
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < in_step; i += kBlock) {
                    const uint64_t kk = staged_key[i];
                    const uint32_t digit = uint32_t(kk >> shift) & mask;
                    const uint32_t dst = step_base[digit] + (i - local_start[digit]);
                    kout[dst] = kk;
                    vout[dst] = staged_val[i];
                }
                __syncthreads();
                for (uint32_t d = threadIdx.x; d < radix; d += kBlock) {
#pragma unroll
                    for (int w = 0; w < kBlock / 32; ++w) warp_count[w][d] = 0;
                }
                __syncthreads();
            }
            // the next pass reads what this one wrote (other threads' stores): made visible by the barrier above
        }
    }
}

}  // namespace

// Sorts keys[0]/vals[0]; *result_buffer receives which of the two buffers holds
// the sorted data.
int sort_passes(int key_bits) { return (key_bits + 8) / 9; }

// Passes of the job-segmented sort for a frame, or 0 when the frame does not qualify on the host's side: it needs
// many jobs (the global passes win on a handful of large ones) and must end in the buffer the global passes end in.
int segmented_sort_passes(int key_bits, int bits_x, int bits_y, size_t n_jobs)
{
    if (n_jobs < 1024) return 0;
    const int need = (bits_x + bits_y + 8) / 9;
    int passes = need;
    if ((passes & 1) != (sort_passes(key_bits) & 1)) ++passes;
    return passes < sort_passes(key_bits) ? passes : 0;
}

void launch_sort(const device_frame &f, cudaStream_t s, int key_bits, int *result_buffer, int seg_passes, int bits_yx, uint32_t n_jobs)
{
    int passes = sort_passes(key_bits);
    int width = (key_bits + passes - 1) / passes;      // <= 9
    if (seg_passes > 0 && f.job_run_begin) {
        launch_pdl(k_job_runs, (n_jobs + kBlock - 1) / kBlock, kBlock, 0, s, f);
        // 3 CTAs per SM: with 4 the ranges in flight (2 buffers each) outgrow the 126 MB L2 and the ping-pong goes to HBM
        // (4.10 ms; 3: 3.96 ms; 2: 4.38 ms on 2048 canvases)
        launch_pdl(k_sort_jobs, kSMs * 3, kBlock, 0, s, f, seg_passes, (bits_yx + seg_passes - 1) / seg_passes);
    }
    int src = 0;
    for (int p = 0; p < passes; ++p) {
        launch_pdl(k_sort_hist, kGrid, kBlock, 0, s, f, src, p * width, width);
        launch_pdl(k_sort_scan, 1 << width, kBlock, 0, s, f);
        launch_pdl(k_sort_scatter, kGrid, kBlock, 0, s, f, src, p * width, width);
        src ^= 1;
    }
    *result_buffer = src;
}

}  // namespace cb200
