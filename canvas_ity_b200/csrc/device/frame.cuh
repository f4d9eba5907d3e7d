// frame.cuh -- device buffers of one frame in flight and the launcher
// prototypes of every pipeline stage (definitions in the .cu files next to it).
#pragma once

#include "common.cuh"

namespace cb200 {

// A polyline/polygon in the shared point pool.  Loop ids are global:
//   [0, n_subpaths)                        K1 output, one per input subpath
//   [n_subpaths, +n_dash_subpaths)         K2 output (dashed strokes only)
//   [stroke_loop_base, +2 * n_sources)     K3 output (two per stroke source)
struct loop_span { uint32_t first, count; };

// A polyline that K3 expands: static ones (un-dashed strokes, loop = subpath id)
// are listed by the host, dashed ones are appended by K2.
struct stroke_src { uint32_t loop, draw_closed; };   // draw | closed << 31
struct dash_item { uint32_t subpath, flags; };       // flags: 1 first of its draw, 2 last of its draw

struct device_frame {
    frame_header *hdr;            // device
    // inputs (uploaded once per frame)
    draw_rec *draws;
    subpath_rec *subpaths;
    unit_rec *units;
    float2 *in_points;                                 // uploaded points, then the glyph-instance region
    glyph_inst_rec *glyph_insts;  uint32_t n_glyph_insts;
    atlas_dev *atlas_table;
    brush_rec *brushes;
    float4 *colors;  float *stops;
    float *dashes;
    dash_item *dash_items; uint32_t n_dash_items;      // subpaths of dashed strokes, draw order
    uint2 *draw_src;                                   // per draw: first stroke source, count
    stroke_src *sources;   uint32_t n_static_sources;
    job_rec *jobs;
    comp_rec *comp;                                    // per job, built by k_job_tiles
    // per tile row of the target: the jobs (in order) whose composite box reaches that row, so the
    // compositor scans ~1/5 of the job table on a picture like the tiger; null = scan the canvas' range
    uint32_t *row_jobs, *row_job_count;  uint32_t row_stride;
    uint32_t *blur_units;                              // [2][n_shadow_jobs + 1] prefix of blur sweep units (x, y)
    uint2 *job_box;  uint32_t *job_te;                 // compact per-job tile box + first tile entry
    // per tile of the target: 1 + the last job that covers the whole tile with an opaque colour (0: none) -- every
    // earlier job and the old pixels are irrelevant there (occlusion culling, k_tile_flags -> k_composite)
    uint32_t *tile_cover;  uint32_t n_target_tiles;
    uint32_t n_opaque_jobs;                            // occlusion-culling candidates in this frame
    int general_compositor;                            // 0 lean, 1 + masks / shadows / clips, 2 lean + small gradients, 3 everything, 4 lean + patterns
    float4 *texels;
    // geometry
    uint32_t *unit_count, *unit_offset;               // n_units + 1
    float2 *pts;           uint32_t cap_pts;
    uint32_t *pt_loop;
    loop_span *loops;      uint32_t cap_loops;
    uint32_t *dash_pts_count, *dash_sub_count, *dash_tail, *dash_pts_off, *dash_sub_off;
    uint32_t *half_count, *half_offset;               // 2 per stroke source (+1)
    uint32_t *half_unit_off, *half_dirty, *stroke_unit_pts;  uint32_t cap_stroke_units;
    uint32_t *half_last, *visit_prev;  uint8_t *visit_close;
    uint32_t cap_sources;
    uint32_t stroke_loop_base;
    // raster
    float4 *pieces;        uint32_t *piece_job;  uint32_t *piece_rows;   // 3 per item
    uint32_t *piece_rlo, *piece_row_off;
    uint32_t cap_items;
    uint32_t *row_runs, *row_piece;  uint32_t cap_rows;   // per (piece,row): run count, piece slot
    uint64_t *keys[2];     float *vals[2];            uint32_t cap_runs;
    float *cumulative;
    uint32_t *long_rows;                               // segment heads too long for one thread
    // shadow working rectangles: loops of shadow jobs that leave the padded canvas (k_edges marks each once
    // and lists it, k_shadow_boxes walks their crossings -- shadow_box.cuh)
    uint32_t *loop_mark;   uint2 *box_loops;           // per loop id; (job, loop id)
    leak_rec *leaks;       uint32_t cap_leaks;         // rows whose coverage residue reaches the right canvas edge
    // tiles
    uint32_t *te_flags, *te_job;  float *te_backdrop;  uint32_t *te_first, *te_mask;  uint32_t cap_tiles;
    float *planes, *planes_tmp;  uint64_t cap_planes;
    uint32_t *shadow_jobs; uint32_t n_shadow_jobs;     // job indices with kind JOB_SHADOW
    int max_shadow_pad, max_shadow_radius, min_shadow_radius;
    // scratch
    uint32_t *partials;                                // several kGrid-sized slices
    uint32_t *sort_hist;
    uint32_t *job_run_begin;                           // [n_jobs + 1] first run of every job in emission order (sort.cu); null: no segmented sort
};

struct canvas_target {
    float4 *fb;            // band_rows x width, linear premultiplied
    int width, height, band_y0, band_rows;
    float **mask_planes;   // device array of plane pointers indexed by slot (slot 0 unused)
    uint32_t n_masks;
    // batches: n_canvases equal canvases stacked vertically in fb, slot_rows (multiple of 32) apart;
    // every job belongs to one canvas and works in that canvas' own coordinates
    int n_canvases, slot_rows;
    const uint2 *canvas_jobs;   // per canvas: first job, job count
    // the frame starts from transparent black: the compositor never loads fb and also writes the
    // tiles no job reaches (a clear folded into the frame instead of a separate 16 B/pixel memset)
    int clear_first;
};

// geometry.cu
void launch_glyphs(const device_frame &f, cudaStream_t s);
void launch_flatten(const device_frame &f, uint32_t n_units, cudaStream_t s);
void launch_dash(const device_frame &f, cudaStream_t s);
void launch_stroke(const device_frame &f, cudaStream_t s);
void launch_join_math(const float *x, uint32_t n, float *acos_out, float *tan_out, cudaStream_t s);   // debug tap
// raster.cu
void launch_raster(const device_frame &f, const canvas_target &t, cudaStream_t s);
// sort.cu
void launch_sort(const device_frame &f, cudaStream_t s, int key_bits, int *result_buffer, int seg_passes, int bits_yx, uint32_t n_jobs);
int segmented_sort_passes(int key_bits, int bits_x, int bits_y, size_t n_jobs);
int sort_passes(int key_bits);
// coverage.cu
void launch_rows(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s);
// composite.cu
void launch_composite(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s);
// shadow.cu
void launch_shadow(const device_frame &f, const canvas_target &t, int sorted_buffer, cudaStream_t s,
                   cudaEvent_t after_raster);
// pixels.cu
// dst / src are tightly packed RGBA8 device buffers of dst_w x dst_h pixels
void launch_readback(const float4 *fb, int width, int band_y0, int band_rows, uint8_t *dst, int dst_w,
                     int dst_h, int x, int y, cudaStream_t s, int bgra = 0);
void launch_upload(float4 *fb, int width, int band_y0, int band_rows, const uint8_t *src, int src_w,
                   int src_h, int x, int y, cudaStream_t s);
void launch_texel_convert(const uint8_t *src, float4 *dst, uint64_t n_texels, cudaStream_t s);
void launch_fill_f32(float *dst, float value, uint64_t n, cudaStream_t s);
// PNG encode (the reference driver's write_png, test/test.cpp:2415-2507) straight from the float framebuffer
struct png_sums { unsigned long long a, b; uint32_t crc, pad; };
size_t png_table_words(int width, int height);
void launch_png_tables(uint32_t *tables, int width, int height, cudaStream_t s);
void launch_png_encode(const float4 *fb, int width, int height, uint8_t *out, const uint32_t *tables, png_sums *acc,
                       uint32_t *row_crc, cudaStream_t s);
// hittest.cu
void launch_hit_test(const float4 *edges, uint32_t n_edges, const float2 *queries, uint32_t n_queries, int2 *acc,
                     uint8_t *inside, cudaStream_t s);

}  // namespace cb200
