// common.cuh -- device-side records and block-level primitives shared by the
// sm_100a kernels of the canvas back end.
//
// Frame pipeline (one pass over ALL draws of a submitted frame; nothing is
// launched per draw -- SURVEY 7.1 "record, don't execute"):
//
//   K1 flatten   cubics  -> polyline points           (geometry.cu; hpp:1331-1524)
//   K2 dash      polylines -> dashed polylines        (geometry.cu; hpp:1858-1934)
//   K3 stroke    polylines -> closed outline loops    (geometry.cu; hpp:1949-2100)
//   K4 raster    loop edges -> signed-area pixel runs (raster.cu;   hpp:2109-2240)
//   K5 sort      runs by (job, y, x), LSD radix       (sort.cu;     hpp:2243)
//   K6 rows      per-scanline running coverage + tile binning (coverage.cu; hpp:2244-2252, 2570)
//   K7 tiles     paint + Porter-Duff per 32x32 tile, draws replayed in order
//                (composite.cu; hpp:2265-2377, 2551-2605, 3057-3099)
//   K9 shadow    alpha plane, separable extended-box blur (shadow.cu; hpp:2395-2539)
//   K10/K11      sRGB+dither readback, RGBA8 upload   (pixels.cu;   hpp:3348-3408)
//
// All element counts that depend on the data live in a device-resident
// frame_header; kernels are launched with fixed grids and read their extents
// from it, so a frame is a fixed launch sequence with no host round trip.
#pragma once

#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "../geom.cuh"
#include "../../../include/canvas_b200.h"

namespace cb200 {

constexpr int kTile = 32;                  // compositor tile is kTile x kTile pixels
constexpr int kBlock = 256;                // threads per CTA for the 1-D work kernels
constexpr int kSMs = 148;                  // B200
constexpr int kGrid = kSMs * 4;            // fixed 1-D grid: 4 CTAs per B200 SM
constexpr float kThreshold = 1.0f / 8160.0f;   // hpp:1289, 2429, 2573
constexpr uint32_t kNoRun = 0xffffffffu;
constexpr int kBlurChunkY = 256;           // rows per chunk of the blur's y sweep (shadow.cu; job_rec::chunk_lo)

// job.kind
enum { JOB_MAIN = 0, JOB_SHADOW = 1, JOB_CLIP = 2 };

// Device copy of one cb200_draw plus what the host precomputed for it.
struct draw_rec {
    uint32_t kind, op;
    uint32_t first_subpath, n_subpaths;
    uint32_t brush, mask_src, mask_dst;
    uint32_t cap, join;
    uint32_t first_dash, n_dash;
    float dash_offset, global_alpha, line_width, miter_limit;
    affine forward, inverse;
    float shadow_color[4];
    float shadow_dx, shadow_dy, shadow_blur;
    uint32_t first_unit, n_units;          // flatten units (start points + cubics)
    float angular;                         // -1 for fills / clips
    uint32_t canvas;
};

struct subpath_rec {
    uint32_t first_point, n_cubics, closed, draw;
    uint32_t first_unit;
};

// One glyph of a text draw: a cached outline (atlas arrays in device memory) placed by `m`; K0
// writes its control points to in_points[first_point ...] (include/canvas_b200.h, cb200_glyph_inst).
struct glyph_inst_rec {
    uint32_t atlas, outline, first_point, reserved;
    affine m;
};
struct atlas_dev {
    const cb200_glyph_outline *outlines;
    const cb200_glyph_seg *segs;
    const float2 *points;
};

// A flatten unit: unit 0 of a subpath re-emits its start point, unit k>0
// flattens cubic k-1.
struct unit_rec { uint32_t subpath, index; };

struct brush_rec {
    uint32_t type, flags, first_color, n_colors;
    float sx, sy, ex, ey, r0, r1;
    uint32_t repetition;
    int32_t width, height;
    uint64_t texel_offset;                 // float4 texels, premultiplied linear
};

// One rasterisation job = one closed-loop set rasterised with one offset into
// one target.  A draw with a shadow owns two jobs (shadow first).
struct job_rec {
    uint32_t draw, kind;
    float off_x, off_y;                    // added to every point (shadow offset + border)
    int32_t pad;                           // target is (W + pad) x (H + pad), hpp:2198
    int32_t border;
    // filled on the device:
    int32_t min_x, min_y, max_x, max_y;    // conservative pixel bounds of the clipped edges
    int32_t run_min_x, run_min_y, run_max_x, run_max_y;   // exact bounds of non-zero runs
    uint32_t first_key;                    // smallest (y << 16 | x) over all runs
    int32_t tx0, ty0, tw, th;              // tile rectangle in raster space (padded for shadows)
    int32_t cx0, cy0, cx1, cy1;            // canvas pixels this job composites into
    uint32_t te_base;                      // first tile entry
    uint32_t first_point;                  // loop points of the draw ...
    uint32_t first_item, n_items;          // ... = K4 work items of this job
    // shadow plane (JOB_SHADOW): working rectangle in padded space + storage
    int32_t left, top, bw, bh;
    // storage: pixel (left + c, top + r) is plane[r * pitch + skew + c]; pitch is a multiple of 32
    // floats and skew puts canvas x = 0 (mod 32) on a 128-byte boundary, so that the raster, both
    // blur sweeps and the compositor move whole aligned 128-byte lines
    int32_t pitch, skew;
    uint64_t plane_offset;
    float w1, w2; int32_t radius;
    // What of the plane this canvas (band) needs, in plane rows (k_job_tiles): the y sweep runs in chunks of
    // kBlurChunkY rows on a grid fixed to the plane's top, each with its own 3 (r + 1) run-in, so a band only sweeps
    // the chunks [chunk_lo, chunk_lo + chunk_n) that hold its rows -- with results bit-identical to a whole-canvas
    // render -- and the raster and the x sweep only feed them: rows [need_r0, need_r1).
    int32_t need_r0, need_r1, chunk_lo, chunk_n;
    uint32_t opaque;                       // host: solid, alpha 1, source_over/copy, unclipped
    uint32_t canvas;                       // batch slot this job draws into
};

// Everything the tile compositor needs to know about one job, one 128-byte line.  The first 48 bytes are what every
// job needs (three 16-byte loads with a warp-uniform address); the rest only shadows, masks and non-solid brushes.
struct comp_rec {
    uint32_t kind, op, flags, brush_type;              // brush_type 0xff: an empty brush paints nothing
    float color[4];                                    // solid colour, or the shadow tint
    float alpha;                                       // global_alpha
    uint32_t mask_src, mask_dst, brush;
    int32_t cx0, cy0, cx1, cy1;                        // canvas pixels a shadow job composites into
    int32_t border, left, top, bw;                     // shadow plane placement: storage origin (left - skew, top), row pitch
    uint32_t plane_lo, plane_hi;                       // plane offset (floats), 64 bit
    uint32_t draw, te_base;
    int32_t tx0, ty0, tw, th;
    uint32_t pad[4];
};
static_assert(sizeof(comp_rec) == 128, "comp_rec must be one 128 B line");
enum { COMP_EVERYWHERE = 1, COMP_OPAQUE = 2 };   // comp_rec.flags
enum { TE_NONEMPTY = 1, TE_COVERED = 2 };        // te_flags
// job_box[j]: tile rectangle the job composites into, 11 bits per coordinate, + kind and flags:
//   x = tx0 | ty0 << 11 | (tx1 & 0x3ff) << 22      y = tx1 >> 10 | ty1 << 1 | kind << 12 | flags
enum { JOBBOX_EVERYWHERE = 1u << 14, JOBBOX_OPAQUE = 1u << 15, JOBBOX_LEAKY = 1u << 16 };

// A scanline of a draw whose signed areas do not add up to zero: the coverage left over after the row's last
// run.  render_main (hpp:2551-2605) walks the runs merged with the clip mask's, whose last run of every row
// sits at the right canvas edge, so a residue of at least 1/8160 is painted all the way to that edge -- beyond
// the draw's own bounding box (float rounding at coordinates in the thousands makes such residues real: tiger at
// 4096^2, one row of draw 102).  k_rows lists them; the compositor consults the list for tiles to the right of a
// leaky job's tile rectangle.
struct leak_rec { uint32_t job; int32_t y; float sum; uint32_t pad; };

struct frame_header {
    // inputs
    uint32_t n_draws, n_subpaths, n_units, n_jobs;
    int32_t width, height, band_y0, band_rows;
    // device-produced counts
    uint32_t n_line_points;                // K1 output
    uint32_t n_dash_points, n_dash_subpaths;
    uint32_t n_sources;                    // stroke sources: static + dashed
    uint32_t n_stroke_points;              // K3 output
    uint32_t n_stroke_units;               // K3 work items (joins + caps)
    uint32_t n_items;                      // K4 work items (job x loop point)
    uint32_t n_row_items;                  // (piece, scanline) pairs
    uint32_t n_runs;
    uint32_t n_tile_entries;
    uint32_t n_long_rows;                  // scanline segments handed to k_rows_long
    uint32_t n_leaks;                      // scanlines of source_over-like draws whose coverage does not return to 0 (leak_rec)
    uint32_t n_box_loops;                  // loops of shadow jobs that cross a side or the top of the padded canvas
    uint32_t max_job_runs;                 // runs of the frame's largest job (k_job_runs; 0: not measured -- sort.cu)
    uint64_t plane_floats;               // storage (pitched)
    uint64_t shadow_working_pixels;      // sum of bw * bh: what the reference blurs (hpp:2426)
    uint32_t overflow;                     // bit set: which capacity was exceeded
    uint32_t sort_bits_x, sort_bits_y, sort_bits;
    unsigned long long composited_pixels, shadow_pixels;
    uint32_t tickets[16];                  // last-block-done counters, zeroed per frame
    // Statistics of the tile compositor (composited pixels), spread over 32 counters on cache lines of their own:
    // a quarter of a million warps adding to ONE word next to `overflow` kept that line busy with atomics while
    // every warp of the kernel starts by reading it (37 % of the kernel's stall samples on an 8192^2 fill).
    alignas(128) unsigned long long composited_slots[32];
};

enum { OVF_POINTS = 1, OVF_LOOPS = 2, OVF_ITEMS = 4, OVF_ROWS = 8, OVF_RUNS = 16, OVF_TILES = 32,
       OVF_PLANES = 64, OVF_DASH = 128, OVF_LEAKS = 256 };

// ---- block-level primitives (kBlock threads) ----------------------------------

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// Programmatic dependent launch: every kernel of a frame is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be placed while the previous
// kernel drains, and starts with grid_dependency_wait() -- nothing written by an earlier kernel is
// read before that returns (it returns once all preceding kernels have completed and flushed).
__device__ __forceinline__ void grid_dependency_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

#ifdef __CUDACC__
// CB200_PDL=0 turns the early placement off (plain stream order) -- an A/B switch for measurements.
inline int pdl_allowed()
{
    static const int allowed = [] { const char *e = getenv("CB200_PDL"); return e ? atoi(e) != 0 : 1; }();
    return allowed;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_allowed();
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

// Exclusive scan of one value per thread across the CTA; returns the exclusive
// prefix and writes the CTA total to `total`.  `smem` needs 33 words.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        uint32_t winc = warp_inclusive_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    uint32_t res = smem[warp] + inc - v;
    total = smem[32];
    __syncthreads();
    return res;
}

// Contiguous slice of [0, n) owned by CTA `b` of `g`: the same mapping is used
// by a count kernel and its emit kernel so their prefix sums line up.
__device__ __forceinline__ void block_slice(uint32_t n, uint32_t &begin, uint32_t &end,
                                            uint32_t &per_thread)
{
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per_thread = (per + blockDim.x - 1) / blockDim.x;            // consecutive items per thread
    per = per_thread * blockDim.x;
    uint64_t b = uint64_t(per) * blockIdx.x;
    begin = b < n ? uint32_t(b) : n;
    end = b + per < n ? uint32_t(b + per) : n;
}

// Slice for kernels that give one WARP per item: whole warps-worth of items per CTA,
// spread over the entire grid.
__device__ __forceinline__ void warp_slice(uint32_t n, uint32_t &begin, uint32_t &end)
{
    const uint32_t warps = blockDim.x >> 5;
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + warps - 1) / warps * warps;
    uint64_t b = uint64_t(per) * blockIdx.x;
    begin = b < n ? uint32_t(b) : n;
    end = b + per < n ? uint32_t(b + per) : n;
}

// Called by every CTA of a count kernel after it wrote partial[blockIdx.x]: the
// last CTA to arrive turns partial[] into exclusive bases and publishes the
// grand total.  Saves a separate scan launch per count/emit pair.
__device__ __forceinline__ void finish_partials(uint32_t *partial, uint32_t *ticket, uint32_t *total_out,
                                                uint32_t *smem)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    uint32_t carry = 0;
    for (uint32_t base = 0; base < gridDim.x; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < gridDim.x ? ((volatile uint32_t *)partial)[i] : 0;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(v, smem, tot);
        if (i < gridDim.x) partial[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) { *total_out = carry; *ticket = 0; }
}

}  // namespace cb200
