/* canvas_b200.h -- C ABI of the B200 rasterise + composite back end.
 *
 * This is the drop-in boundary for canvas_ity's data-parallel hot path.  The
 * reference (src/canvas_ity.hpp, "hpp" below) has no FFI of its own: its seam is
 * the private call chain behind the draw entry points
 *     fill()            hpp:3044   path_to_lines + render_main
 *     stroke()          hpp:3050   path_to_lines + stroke_lines + render_main
 *     clip()            hpp:3057   path_to_lines + lines_to_runs + mask product
 *     fill_rectangle()  hpp:3155 / stroke_rectangle() hpp:3174 / draw_image() hpp:3313
 *     fill_text()       hpp:3276 / stroke_text() hpp:3286
 *     get_image_data()  hpp:3348 / put_image_data()  hpp:3383
 * Host C++ keeps recording state and device-space cubic paths exactly as the
 * reference does (hpp:2866-3042) and lowers every draw call to one cb200_draw
 * record; cb200_submit() replaces everything the reference does below those
 * entry points (hpp:1331-2605) with sm_100a kernels.
 *
 * Plain C: pointers + sizes only, no C++/torch types.  Every function returns
 * 0 on success or a negative cb200_status; nothing throws across this ABI.
 * There is no CPU fallback: without a CUDA device every call fails with
 * CB200_ERR_NO_DEVICE.
 */
#ifndef CANVAS_B200_H
#define CANVAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_ABI_VERSION 2

typedef enum cb200_status {
    CB200_OK = 0,
    CB200_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: there is no CPU path */
    CB200_ERR_BAD_ARG = -2,
    CB200_ERR_CUDA = -3,        /* see cb200_last_error() */
    CB200_ERR_OOM = -4,
    CB200_ERR_OVERFLOW = -5     /* a device work queue overflowed even after regrowth */
} cb200_status;

/* ---- lowered draw records (the data contract, hpp:159-175) ---------------- */

/* cb200_draw.kind */
enum { CB200_FILL = 0, CB200_STROKE = 1, CB200_CLIP = 2 };
/* cb200_brush.type == paint_brush::types, hpp:162 */
enum { CB200_BRUSH_COLOR = 0, CB200_BRUSH_LINEAR = 1, CB200_BRUSH_RADIAL = 2,
       CB200_BRUSH_PATTERN = 3 };
/* cb200_brush.flags */
enum { CB200_BRUSH_CLAMP = 1 };   /* draw_image addressing, hpp:2306/2319 */

/* One subpath: a start point followed by n_cubics * 3 control points, all in
 * device space (hpp:2878-2882).  Lines are degenerate cubics (hpp:2913-2916). */
typedef struct cb200_subpath {
    uint32_t first_point;   /* index into cb200_frame.points (xy pairs) */
    uint32_t n_cubics;      /* subpath owns 1 + 3 * n_cubics points */
    uint32_t closed;        /* subpath_data.closed, hpp:169 */
    uint32_t instanced;     /* 1: first_point counts from the start of the frame's glyph-instance
                               region (the points the device expands from cb200_frame.glyphs),
                               which follows points[n_points - 1]; 0: an uploaded point */
} cb200_subpath;

/* paint_brush (hpp:162-165) flattened.  Colours: solid = 1 premultiplied
 * linear rgba; gradients = n_colors unpremultiplied linear stops (hpp:2834);
 * pattern = an entry of cb200_frame.images. */
typedef struct cb200_brush {
    uint32_t type;
    uint32_t flags;
    uint32_t first_color;   /* index into colors[] (rgba quads) and stops[] */
    uint32_t n_colors;
    float    start[2], end[2];
    float    start_radius, end_radius;
    uint32_t image;         /* pattern: index into images[] */
    uint32_t repetition;    /* repetition_style bit set, hpp:153 / 2278-2282 */
} cb200_brush;

/* A straight-alpha sRGB RGBA8 image (what set_pattern/draw_image receive,
 * hpp:2839/3313); converted on the device to premultiplied linear float. */
typedef struct cb200_image {
    uint64_t texel_offset;  /* byte offset into cb200_frame.texels, tightly packed rows */
    int32_t  width, height;
} cb200_image;

/* State snapshot of one fill/stroke/clip (everything render_main, render_shadow,
 * stroke_lines and dash_lines read, hpp:1858-2605). */
typedef struct cb200_draw {
    uint32_t kind;
    uint32_t op;            /* composite_operation value = 4-bit mix program, hpp:146-149 */
    uint32_t first_subpath, n_subpaths;
    uint32_t brush;
    uint32_t mask_src;      /* clip-mask slot that gates this draw; 0 = whole canvas */
    uint32_t mask_dst;      /* CB200_CLIP: slot that receives coverage * mask_src */
    uint32_t cap, join;     /* cap_style / join_style, hpp:150-151 */
    uint32_t first_dash, n_dash;
    float    dash_offset;
    float    global_alpha;
    float    line_width, miter_limit;
    float    forward[6], inverse[6];   /* affine_matrix a..f, hpp:161 */
    float    shadow_color[4];          /* premultiplied linear, hpp:2726 */
    float    shadow_offset_x, shadow_offset_y, shadow_blur;
    uint32_t reserved;
} cb200_draw;

/* ---- glyph outline cache (device-side text; fill_text hpp:3276, stroke_text hpp:3286) ----
 *
 * The reference re-parses a glyph's TrueType outline on every draw (add_glyph, hpp:1533-1696,
 * text_to_lines hpp:1793-1846).  Here the front end parses each glyph ONCE into a transform-
 * independent outline -- font-unit points plus a list of pieces that refer to them -- which the
 * back end keeps in device memory; a text draw then uploads one cb200_glyph_inst (outline id +
 * 2x3 matrix) per glyph and a kernel writes its device-space cubics straight into the frame's
 * point pool.  The arithmetic per piece is exactly what the host lowering does (same products,
 * same order, no FMA), so an instanced glyph is bit-identical to an uploaded one.
 *
 * End points are symbolic pairs (a, b): the transformed outline point P[a] when a == b, else
 * mix(P[a], P[b], 0.5), the on-curve point TrueType implies between two off-curve points.
 * A curved piece is the quadratic (from, P[ctrl], to), degree-elevated to the cubic
 * (mix(from, C, 2/3), mix(to, C, 2/3), to); a straight piece is the cubic (from, to, to). */
enum { CB200_SEG_LINE = 1,     /* straight piece; ctrl unused */
       CB200_SEG_FIRST = 2 };  /* first piece of a contour: `from` is also stored at out - 1 */
typedef struct cb200_glyph_seg {
    uint16_t from_a, from_b;
    uint16_t to_a, to_b;
    uint16_t ctrl;
    uint16_t flags;
    uint32_t out;           /* the piece's three control points land at out, out + 1, out + 2,
                               counted from the instance's first point */
} cb200_glyph_seg;

typedef struct cb200_glyph_outline {
    uint32_t first_point, n_points;   /* into cb200_glyph_atlas.points */
    uint32_t first_seg, n_segs;       /* into cb200_glyph_atlas.segs, contour by contour */
    uint32_t n_contours;              /* an instance is n_contours closed subpaths ... */
    uint32_t out_points;              /* ... of out_points points in all (n_contours + 3 * n_segs) */
} cb200_glyph_outline;

/* All outlines cached for one font so far.  Append-only: entries never change once published and
 * ids are never reused, so a device copy is extended, never invalidated; the arrays a snapshot
 * points at stay valid for the life of the process. */
typedef struct cb200_glyph_atlas {
    uint64_t id;
    const cb200_glyph_outline *outlines;  uint32_t n_outlines;
    const cb200_glyph_seg     *segs;      uint32_t n_segs;
    const float               *points;    uint32_t n_points;   /* xy pairs, font units */
} cb200_glyph_atlas;

typedef struct cb200_glyph_inst {
    uint32_t atlas;         /* index into cb200_frame.atlases */
    uint32_t outline;
    uint32_t first_point;   /* first output point, counted from the start of the instance region */
    float    m[6];          /* font units -> device space (affine_matrix a..f) */
} cb200_glyph_inst;

typedef struct cb200_frame {
    const cb200_draw    *draws;     uint32_t n_draws;
    const cb200_subpath *subpaths;  uint32_t n_subpaths;
    const float         *points;    uint32_t n_points;    /* xy pairs */
    const cb200_brush   *brushes;   uint32_t n_brushes;
    const float         *colors;    /* rgba quads */
    const float         *stops;     uint32_t n_colors;
    const float         *dashes;    uint32_t n_dashes;
    const cb200_image   *images;    uint32_t n_images;
    const uint8_t       *texels;    uint64_t texel_bytes;
    /* glyph instances (may all be zero): expanded on the device into n_glyph_points points that
     * instanced subpaths refer to */
    const cb200_glyph_atlas *atlases;  uint32_t n_atlases;
    const cb200_glyph_inst  *glyphs;   uint32_t n_glyphs;
    uint32_t n_glyph_points;
} cb200_frame;

/* ---- canvases --------------------------------------------------------------- */

typedef struct cb200_canvas cb200_canvas;

/* New canvas, all pixels transparent black, mask slot 0 = everything visible
 * (ctor, hpp:2607-2644).  `device` is a CUDA ordinal. */
int cb200_canvas_create(int width, int height, int device, cb200_canvas **out);

/* A canvas that owns only scanlines [band_y0, band_y0 + band_rows) of a
 * width x height image; geometry stays in full-image coordinates. */
int cb200_canvas_create_band(int width, int height, int band_y0, int band_rows,
                             int device, cb200_canvas **out);

void cb200_canvas_destroy(cb200_canvas *canvas);

/* A batch of `n_canvases` independent, equally sized canvases that are rendered
 * together: one upload and one launch sequence for all of them (the reference has
 * no shared state between canvases, hpp:73-74, so they are one big data-parallel
 * job).  The result is a cb200_canvas for destroy/sync/stats; submit and reads go
 * through the cb200_batch_* calls, frames carry their canvas index (ascending). */
int cb200_batch_create(int n_canvases, int width, int height, int device, cb200_canvas **out);
int cb200_batch_submit(cb200_canvas *batch, const cb200_frame *const *frames,
                       const uint32_t *canvas_index, uint32_t n_frames);
int cb200_batch_read_rgba8(cb200_canvas *batch, uint32_t canvas, uint8_t *dst, int width,
                           int height, int stride, int x, int y);
int cb200_batch_read_f32(cb200_canvas *batch, uint32_t canvas, float *dst);
/* put_image_data into canvas `canvas` of a batch (hpp:3383-3408), ordered after the frames submitted so far. */
int cb200_batch_write_rgba8(cb200_canvas *batch, uint32_t canvas, const uint8_t *src, int width, int height,
                            int stride, int x, int y);
/* Batch form of cb200_masks_keep: the clip-mask planes still reachable are those of the listed (canvas, that
 * canvas' own slot) pairs; every other plane of the batch is freed. */
int cb200_batch_masks_keep(cb200_canvas *batch, const uint32_t *canvas, const uint32_t *local_slot, uint32_t n);

/* Run `frame` (draws in order) on the canvas' stream.  Asynchronous: returns
 * once the frame is copied to pinned staging and the kernels are enqueued. */
int cb200_submit(cb200_canvas *canvas, const cb200_frame *frame);

/* Upload a frame once, then replay it with no host<->device traffic
 * (device-resident timing; the framebuffer is cleared first when `clear`). */
int cb200_frame_upload(cb200_canvas *canvas, const cb200_frame *frame);
int cb200_frame_replay(cb200_canvas *canvas, int clear);
/* Keep the frame that was submitted last (cb200_submit or cb200_batch_submit) as the resident frame: waits for it to
 * complete, then cb200_frame_replay re-runs it.  This is how a whole batch is replayed (its frame is assembled from the
 * members' frames at submission, there is no single cb200_frame to upload). */
int cb200_frame_keep(cb200_canvas *canvas);

/* Replays of a resident frame that has completed once are issued as ONE CUDA graph launch (the frame's
 * ~36 kernels with their programmatic-dependent-launch edges, header restore and readback) instead
 * of ~45 stream calls: with several canvases replaying concurrently the host's enqueue cost, not the
 * GPU, is what bounds frames/s.  On by default; off = plain stream launches (per-frame compositor
 * events for cb200_timer_end, per-stage events).  Per-stage timing also disables it. */
int cb200_set_graph_replay(cb200_canvas *canvas, int on);

/* Wait for everything enqueued on the canvas' stream. */
int cb200_sync(cb200_canvas *canvas);

/* get_image_data (hpp:3348): sRGB + 4x4 ordered dither to straight-alpha RGBA8.
 * (x, y) is the canvas-space origin of the destination; out-of-canvas = 0. */
int cb200_read_rgba8(cb200_canvas *canvas, uint8_t *dst, int width, int height,
                     int stride, int x, int y);
/* Page-locked host memory for image buffers: cb200_read_rgba8 / cb200_write_rgba8 copy
 * straight between the device and such a buffer (no staging copy).  Any other host
 * pointer works too, through a chunked, pipelined staging copy. */
void *cb200_host_alloc(size_t bytes);
void cb200_host_free(void *ptr);
/* put_image_data (hpp:3383). */
int cb200_write_rgba8(cb200_canvas *canvas, const uint8_t *src, int width,
                      int height, int stride, int x, int y);
/* Linear premultiplied float RGBA framebuffer, band_rows * width * 4 floats. */
int cb200_read_f32(cb200_canvas *canvas, float *dst);
/* Clip-mask slot as dense visibility, band_rows * width floats. */
int cb200_read_mask(cb200_canvas *canvas, uint32_t slot, float *dst);
/* Clear to transparent black and forget all mask slots.  The clear itself is deferred: the next
 * frame absorbs it (its compositor starts from transparent black), any other pixel access applies it. */
int cb200_clear(cb200_canvas *canvas);
/* Clip-mask slots are immutable once written (clip() always writes a fresh
 * slot).  The host tells the back end which slots are still reachable (current
 * mask + save stack); every other plane is freed. */
int cb200_masks_keep(cb200_canvas *canvas, const uint32_t *slots, uint32_t n);

/* cb200_read_rgba8 with the red and blue channels exchanged on the device (the order TGA and BMP
 * files store; demos/tiger/tiger.cpp:4339 swaps them on the CPU after get_image_data). */
int cb200_read_bgra8(cb200_canvas *canvas, uint8_t *dst, int width, int height, int stride, int x, int y);
/* The linear premultiplied float4 framebuffer itself, in device memory (rows x width float4, rows
 * = the band's rows, or all stacked slots of a batch): for zero-copy views from CUDA code or
 * torch (__cuda_array_interface__).  Waits for queued work; valid until the next frame. */
int cb200_framebuffer_device(cb200_canvas *canvas, void **device_ptr, int *rows, int *width);

/* The whole PNG file the reference's driver writes after get_image_data (write_png,
 * test/test.cpp:2415-2507: IHDR + sRGB + one IDAT of stored deflate blocks, one per row, + IEND),
 * produced on the device straight from the float framebuffer: one kernel does the sRGB + dither
 * conversion, lays the bytes out as the file and folds the Adler-32 and CRC-32 in, so no CPU pass
 * over the pixels follows the readback.  Byte-identical to write_png() over get_image_data()'s
 * output.  dst == NULL: only *bytes (the file size, 76 + height * (6 + 4 * width)) is returned.
 * Whole canvases only (no bands / batches); 4 * width + 1 must fit a stored block (width <= 16383). */
int cb200_encode_png(cb200_canvas *canvas, uint8_t *dst, size_t capacity, size_t *bytes);

/* Readback without the host copy: runs the sRGB/dither kernel into a device
 * buffer owned by the canvas and returns its device pointer (for NCCL gathers
 * and device-resident timing). */
int cb200_read_rgba8_device(cb200_canvas *canvas, void **device_ptr);
/* Same kernel, but into a caller-owned DEVICE buffer of width*height*4 bytes (a
 * torch tensor, an NCCL send buffer): asynchronous on the canvas stream, follow
 * with cb200_sync() before handing the buffer to another stream. */
int cb200_read_rgba8_into(cb200_canvas *canvas, void *device_dst, int width, int height, int x, int y);

/* Bulk is_point_in_path (hpp:3101-3132): the reference answers one point per call by walking every
 * edge of the flattened path; this tests n_queries points against the same edges in one launch.
 * `edges`: n_edges x (from.x, from.y, to.x, to.y) in device space -- the flattened current path
 * (path_to_lines, hpp:1495-1524) with every subpath's closing edge included, which is what the
 * front end's cv_path_edges() returns.  `queries`: n_queries xy pairs (device space, like the
 * reference's arguments).  inside[i] = 1 when point i is inside under the non-zero winding rule or
 * exactly on an edge, else 0 -- the reference's bool.  `kernel_ms` (optional) receives the CUDA-event
 * time of the kernels alone.  Synchronous (it returns the answer), like the call it replaces. */
int cb200_hit_test(cb200_canvas *canvas, const float *edges, uint32_t n_edges, const float *queries,
                   uint32_t n_queries, uint8_t *inside, float *kernel_ms);

/* ---- introspection / measurement ------------------------------------------- */

typedef struct cb200_stats {
    uint64_t draws, cubics, line_points, edges, raw_runs, tile_entries,
             composited_pixels, shadow_pixels, kernel_launches;
    float    last_frame_ms;        /* CUDA events on the canvas stream */
    float    composite_ms;         /* the tile compositor only */
    float    raster_ms, sort_ms, geometry_ms, readback_ms;
    float    coverage_ms;          /* scanline walk: running sums + tile entries (after the sort) */
    float    shadow_raster_ms;     /* coverage * alpha into the shadow planes */
    float    blur_ms;              /* both blur sweeps */
    float    png_ms;               /* cb200_encode_png: convert + lay out + checksums, device only */
    uint32_t graph_replays;        /* cb200_frame_replay calls that ran as one CUDA graph launch */
} cb200_stats;

int cb200_get_stats(cb200_canvas *canvas, cb200_stats *out);
/* Per-stage CUDA events (geometry/raster/sort/coverage/shadow) sit between kernels and cost a few
 * microseconds per frame; off: only last_frame_ms and composite_ms are measured.  Default on. */
int cb200_set_stage_timing(cb200_canvas *canvas, int on);
/* Device-side stopwatch on the canvas stream, for timing a run of frames without a host round
 * trip per frame: _begin records an event, _end records another, waits for it and returns the
 * elapsed milliseconds plus the summed duration of the tile compositor over the frames in
 * between (the most recent 256 of them that were issued as stream launches -- graph replays carry no
 * per-frame events; *composite_frames says how many were summed). */
int cb200_timer_begin(cb200_canvas *canvas);
int cb200_timer_end(cb200_canvas *canvas, float *elapsed_ms, float *composite_ms, uint32_t *composite_frames);
/* The canvas' CUDA stream (a cudaStream_t): everything a canvas does is ordered on it, so a caller that chains its
 * own device work -- an NCCL gather of finished bands, say -- enqueues it there (or makes its stream wait on an
 * event recorded there) instead of synchronising with the host. */
void *cb200_stream(cb200_canvas *canvas);
/* Record the end event now without waiting (cb200_timer_end then only waits and reads): with several canvases,
 * stop all of them first so that no end event is recorded late because the host was waiting on another canvas. */
int cb200_timer_stop(cb200_canvas *canvas);
/* Milliseconds from `from`'s _begin event to `to`'s _end event (both recorded, same device): with several canvases
 * replaying concurrently the timed region is the earliest begin to the latest end, the maximum of this over all
 * pairs. */
int cb200_timer_between(cb200_canvas *from, cb200_canvas *to, float *elapsed_ms);

/* Debug taps used by the parity tests: intermediate buffers of the last frame.
 * Each returns the element count (>= 0) and copies at most `capacity`. */
int64_t cb200_debug_lines(cb200_canvas *canvas, float *xy_pairs, uint32_t *job_of_edge,
                          int64_t capacity);       /* edges: 4 floats each */
int64_t cb200_debug_runs(cb200_canvas *canvas, uint64_t *keys, float *cumulative,
                         int64_t capacity);        /* sorted (job,y,x) runs */
/* Host build of what the device enters into a shadow's run bounding box (render_shadow hpp:2409-2419 over the
 * polygon clip of hpp:2208-2229) for the closed loop xy[0 .. n), offset by (off_x, off_y) into a padded
 * canvas of padded_w x padded_h: csrc/device/edge_clip.cuh compiled for the host.  box5 = (min_x, max_x,
 * min_y, max_y, key of the loop's first run y << 16 | x or -1); max < 0 when nothing.  Touches no device:
 * the CPU test checks the rule against the reference's clip. */
void cb200_debug_shadow_box(const float *xy, uint32_t n, float off_x, float off_y, int padded_w, int padded_h,
                            int *box5);

/* Host build of the device's scan conversion (csrc/device/edge_clip.cuh) of the closed loop xy[0 .. n): every
 * signed-area run as run_xy[2 i] = x, run_xy[2 i + 1] = y, run_delta[i], unsorted.  Returns the count and copies at
 * most `capacity`.  Touches no device. */
int64_t cb200_debug_loop_runs(const float *xy, uint32_t n, float off_x, float off_y, int padded_w, int padded_h,
                              int32_t *run_xy, float *run_delta, int64_t capacity);

/* The rounded join's acosf / tanf (hpp:1995-1997) as the kernels compute them (csrc/geom.cuh: glibc's fdlibm
 * algorithms restated), over x[0 .. n): acos_out[i] = acosf(x[i]) for |x| <= 1, tan_out[i] = tanf(x[i]) for
 * 0 <= x <= pi/4.  on_device = 0 runs the host build of the same code, 1 a kernel. */
int cb200_debug_join_math(const float *x, uint32_t n, float *acos_out, float *tan_out, int on_device);

const char *cb200_last_error(void);
int cb200_abi_version(void);
/* sizeof of the ABI records, for bindings that mirror them: 0 draw, 1 subpath, 2 brush, 3 image, 4 frame,
 * 5 glyph_seg, 6 glyph_outline, 7 glyph_atlas, 8 glyph_inst */
int cb200_struct_size(int which);
int cb200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CANVAS_B200_H */
