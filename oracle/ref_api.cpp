// ref_api.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The unmodified reference (REFERENCE_HPP = /root/reference/src/canvas_ity.hpp,
// compiled from where it lies; never copied into this repo) behind the flat C
// API of include/canvas_b200_api.h.  Built by oracle/Makefile into
// oracle/_ref/libcanvas_ref.so with -ffp-contract=off (bit-identical to -O0,
// SURVEY 7.4) and -fno-access-control so cv_read_f32 can reach the private float
// `bitmap` (reference :1171).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the result.
#define CANVAS_ITY_IMPLEMENTATION
#include REFERENCE_HPP

#include "../canvas_ity_b200/csrc/host/script.hpp"
#include "../include/canvas_b200_api.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct ns_tag {
    typedef canvas_ity::composite_operation composite_operation;
    typedef canvas_ity::cap_style cap_style;
    typedef canvas_ity::join_style join_style;
    typedef canvas_ity::brush_type brush_type;
    typedef canvas_ity::repetition_style repetition_style;
    typedef canvas_ity::align_style align_style;
    typedef canvas_ity::baseline_style baseline_style;
};
canvas_ity::canvas *ref(cv_canvas *c) { return reinterpret_cast<canvas_ity::canvas *>(c); }
}

extern "C" {

cv_canvas *cv_create(int width, int height)
{
    return reinterpret_cast<cv_canvas *>(new canvas_ity::canvas(width, height));
}
cv_canvas *cv_create_band(int, int, int, int, int) { return nullptr; }
cv_canvas *cv_create_tapped(int, int, cv_frame_fn, cv_read_fn, cv_write_fn, void *) { return nullptr; }
void cv_destroy(cv_canvas *c) { delete ref(c); }

long cv_run_script(cv_canvas *canvas, const uint8_t *script, size_t bytes, uint32_t *queries,
                   int query_capacity, int *n_queries)
{
    std::vector<cb200_script::query_result> q;
    long n = cb200_script::run_script<canvas_ity::canvas, ns_tag>(*ref(canvas), script, bytes, &q);
    if (n_queries) *n_queries = int(q.size());
    for (int i = 0; queries && i < query_capacity && i < int(q.size()); ++i) {
        queries[i * 4 + 0] = q[size_t(i)].code;
        queries[i * 4 + 1] = q[size_t(i)].got_bits;
        queries[i * 4 + 2] = q[size_t(i)].recorded_bits;
        queries[i * 4 + 3] = 0;
    }
    return n;
}

int cv_get_image_data(cv_canvas *c, uint8_t *image, int w, int h, int stride, int x, int y)
{
    ref(c)->get_image_data(image, w, h, stride, x, y);
    return 0;
}
int cv_put_image_data(cv_canvas *c, const uint8_t *image, int w, int h, int stride, int x, int y)
{
    ref(c)->put_image_data(image, w, h, stride, x, y);
    return 0;
}
int cv_is_point_in_path(cv_canvas *c, float x, float y) { return ref(c)->is_point_in_path(x, y); }
float cv_measure_text(cv_canvas *c, const char *text) { return ref(c)->measure_text(text); }
// the reference has no bulk query: n calls of its own is_point_in_path (hpp:3101-3132)
int cv_points_in_path(cv_canvas *c, const float *xy, int n, uint8_t *inside)
{
    for (int i = 0; i < n; ++i) inside[i] = ref(c)->is_point_in_path(xy[2 * i], xy[2 * i + 1]) ? 1 : 0;
    return 0;
}
long cv_path_edges(cv_canvas *, float *, long) { return -1; }
// what demos/tiger/tiger.cpp:4333-4345 does after rendering: get_image_data, swap R and B on the CPU, write
int cv_write_tga(cv_canvas *c, const char *path)
{
    canvas_ity::canvas *r = ref(c);
    const int w = r->size_x, h = r->size_y;
    std::vector<unsigned char> image(size_t(w) * size_t(h) * 4);
    r->get_image_data(image.data(), w, h, 4 * w, 0, 0);
    for (size_t i = 0; i < image.size(); i += 4) { unsigned char t = image[i]; image[i] = image[i + 2]; image[i + 2] = t; }
    const unsigned char header[18] = { 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, (unsigned char)(w & 255), (unsigned char)(w >> 8),
                                       (unsigned char)(h & 255), (unsigned char)(h >> 8), 32, 40 };
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    fwrite(header, 1, sizeof header, f);
    fwrite(image.data(), 1, image.size(), f);
    fclose(f);
    return 0;
}
// the reference driver's own PNG writer (test/test.cpp:2415-2507), compiled in ref_png.cpp
void ref_write_png(const char *path, const unsigned char *rgba, int width, int height);
int cv_write_png(cv_canvas *c, const char *path)
{
    canvas_ity::canvas *r = ref(c);
    const int w = r->size_x, h = r->size_y;
    std::vector<unsigned char> image(size_t(w) * size_t(h) * 4);
    r->get_image_data(image.data(), w, h, 4 * w, 0, 0);
    ref_write_png(path, image.data(), w, h);
    return 0;
}
int cv_set_text_instancing(cv_canvas *, int) { return 0; }
int cv_flush(cv_canvas *) { return 0; }
// batches are a back-end concept: the reference renders canvases one by one
cv_batch *cv_batch_create(int, int, int, int) { return nullptr; }
cv_canvas *cv_batch_canvas(cv_batch *, int) { return nullptr; }
int cv_batch_flush(cv_batch *) { return -1; }
int cv_batch_get_image_data(cv_batch *, int, uint8_t *, int, int, int, int, int) { return -1; }
int cv_batch_read_f32(cv_batch *, int, float *) { return -1; }
cb200_canvas *cv_batch_device(cv_batch *) { return nullptr; }
void cv_batch_destroy(cv_batch *) {}

int cv_read_f32(cv_canvas *c, float *dst)
{
    canvas_ity::canvas *r = ref(c);
    memcpy(dst, r->bitmap, sizeof(float) * 4 * size_t(r->size_x) * size_t(r->size_y));
    return 0;
}

cb200_canvas *cv_device(cv_canvas *) { return nullptr; }
const char *cv_last_error(void) { return ""; }
const char *cv_backend_name(void) { return "reference"; }

}  // extern "C"
