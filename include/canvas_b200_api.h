/* canvas_b200_api.h -- flat C entry points over the drop-in `canvas_ity::canvas`
 * front end (include/canvas_ity.hpp), for ctypes / cgo / JNI style bindings.
 *
 * A binding builds a "canvas script" (canvas_ity_b200/csrc/host/script.hpp: one
 * opcode per reference API method, src/canvas_ity.hpp:194-1148) and runs it with
 * cv_run_script(); synchronous queries have direct entry points.  The same
 * signatures are implemented over the unmodified reference class in
 * oracle/ref_api.cpp (-> oracle/_ref/libcanvas_ref.so) so one test body can
 * drive both.
 */
#ifndef CANVAS_B200_API_H
#define CANVAS_B200_API_H

#include <stddef.h>
#include <stdint.h>

#include "canvas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cv_canvas cv_canvas;

/* canvas(width, height), reference :194.  NULL + cv_last_error() on failure
 * (no CUDA device: there is no CPU fallback). */
cv_canvas *cv_create(int width, int height);
/* One scanline band [band_y0, band_y0+band_rows) of a width x height image on
 * CUDA device `device` (multi-GPU sharding of one large canvas). */
cv_canvas *cv_create_band(int width, int height, int device, int band_y0, int band_rows);
/* Lowering only: no device is touched, every flushed frame is handed to
 * `on_frame`, pixel traffic to `on_read` / `on_write`.  Used by the parity tests
 * to feed the identical lowered frames to the oracle. */
typedef void (*cv_frame_fn)(void *user, const cb200_frame *frame);
typedef void (*cv_read_fn)(void *user, uint8_t *dst, int w, int h, int stride, int x, int y);
typedef void (*cv_write_fn)(void *user, const uint8_t *src, int w, int h, int stride, int x, int y);
cv_canvas *cv_create_tapped(int width, int height, cv_frame_fn on_frame, cv_read_fn on_read,
                            cv_write_fn on_write, void *user);
void cv_destroy(cv_canvas *canvas);

/* Replay a canvas script.  Returns the number of calls executed (< 0: malformed).
 * Query results met on the way (is_point_in_path, measure_text, set_font,
 * get_image_data) are written as 4 uint32 each {opcode, got bits, recorded bits, 0}
 * up to `query_capacity` entries; *n_queries receives how many there were. */
long cv_run_script(cv_canvas *canvas, const uint8_t *script, size_t bytes,
                   uint32_t *queries, int query_capacity, int *n_queries);

/* Synchronous calls of the reference API. */
int   cv_get_image_data(cv_canvas *canvas, uint8_t *image, int width, int height,
                        int stride, int x, int y);                 /* :1092 */
int   cv_put_image_data(cv_canvas *canvas, const uint8_t *image, int width, int height,
                        int stride, int x, int y);                 /* :1123 */
int   cv_is_point_in_path(cv_canvas *canvas, float x, float y);    /* :843 */
float cv_measure_text(cv_canvas *canvas, const char *text);        /* :1025 */

/* Bulk hit testing: is_point_in_path (:843, hpp:3101-3132) for n points at once against the canvas'
 * current path.  xy: n device-space pairs; inside[i] = the reference's bool for point i.  On the
 * B200 build this is one cb200_hit_test launch (a tapped canvas has no device: CB200_ERR_NO_DEVICE,
 * there is no CPU path); on the reference build it is the reference's own loop, n calls. */
int cv_points_in_path(cv_canvas *canvas, const float *xy, int n, uint8_t *inside);
/* The flattened current path as edges (from.x, from.y, to.x, to.y), closing edges included --
 * what cv_points_in_path hands to cb200_hit_test.  Copies at most `capacity` edges, returns the
 * edge count.  B200 build only (returns -1 on the reference build). */
long cv_path_edges(cv_canvas *canvas, float *edges, long capacity);

/* The file demos/tiger/tiger.cpp:4333-4345 writes after get_image_data: an uncompressed 32-bit TGA
 * (18-byte header, top-down rows, BGRA).  Here the channel swap happens in the readback kernel
 * (cb200_read_bgra8) and the rows go from the pinned staging buffer to the file. */
int cv_write_tga(cv_canvas *canvas, const char *path);

/* The file the reference's test driver writes (write_png, test/test.cpp:2415-2507).  On the B200
 * build the whole file image -- pixels, stored-deflate framing, Adler-32, CRC-32 -- comes off the
 * device (cb200_encode_png) and goes to disk as is; on the reference build it is get_image_data
 * followed by the reference's own write_png. */
int cv_write_png(cv_canvas *canvas, const char *path);

/* A batch of n independent width x height canvases rendered together on one GPU
 * (cb200_batch_*): cv_batch_canvas(i) is an ordinary front-end canvas whose draws
 * are queued in the batch; cv_batch_flush() lowers and submits all of them in one
 * device frame.  get_image_data on a member canvas flushes the whole batch. */
typedef struct cv_batch cv_batch;
cv_batch *cv_batch_create(int n_canvases, int width, int height, int device);
cv_canvas *cv_batch_canvas(cv_batch *batch, int index);
int cv_batch_flush(cv_batch *batch);
int cv_batch_get_image_data(cv_batch *batch, int index, uint8_t *image, int width, int height,
                            int stride, int x, int y);
int cv_batch_read_f32(cv_batch *batch, int index, float *dst);
cb200_canvas *cv_batch_device(cv_batch *batch);
void cv_batch_destroy(cv_batch *batch);

/* Text draws normally upload one glyph instance (cached outline id + matrix) per glyph and the
 * device expands it (cb200_glyph_inst); on = 0 makes the front end lower glyph outlines to path
 * points on the host instead, as the reference does per draw (hpp:1533-1696).  Both give the same
 * pixels bit for bit -- the switch exists for the A/B parity tests and for measuring the difference.
 * No-op on the reference build. */
int cv_set_text_instancing(cv_canvas *canvas, int on);

/* Flush queued draws (no-op on the reference build). */
int cv_flush(cv_canvas *canvas);
/* Linear premultiplied float framebuffer, rows * width * 4 floats. */
int cv_read_f32(cv_canvas *canvas, float *dst);
/* The device canvas behind a front-end canvas (NULL for tapped / reference). */
cb200_canvas *cv_device(cv_canvas *canvas);
const char *cv_last_error(void);
/* "b200" or "reference". */
const char *cv_backend_name(void);

#ifdef __cplusplus
}
#endif
#endif /* CANVAS_B200_API_H */
