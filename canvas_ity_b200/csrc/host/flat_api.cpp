// flat_api.cpp -- extern "C" entry points of include/canvas_b200_api.h over the
// recording front end (canvas_front.cpp).
#include "front_state.hpp"
#include "script.hpp"

#include "../../../include/canvas_b200_api.h"

#include <new>
#include <stdexcept>
#include <string>

namespace {
thread_local std::string g_api_error;

struct ns_tag {
    typedef canvas_ity::composite_operation composite_operation;
    typedef canvas_ity::cap_style cap_style;
    typedef canvas_ity::join_style join_style;
    typedef canvas_ity::brush_type brush_type;
    typedef canvas_ity::repetition_style repetition_style;
    typedef canvas_ity::align_style align_style;
    typedef canvas_ity::baseline_style baseline_style;
};

canvas_ity::canvas *front(cv_canvas *c) { return reinterpret_cast<canvas_ity::canvas *>(c); }
}

extern "C" {

cv_canvas *cv_create(int width, int height)
{
    try {
        return reinterpret_cast<cv_canvas *>(new canvas_ity::canvas(width, height));
    } catch (const std::exception &e) {
        g_api_error = e.what();
        return nullptr;
    }
}

cv_canvas *cv_create_band(int width, int height, int device, int band_y0, int band_rows)
{
    if (device < 0) { g_api_error = "cv_create_band: device must be >= 0"; return nullptr; }
    try {
        return reinterpret_cast<cv_canvas *>(
            new canvas_ity::canvas(width, height, device, band_y0, band_rows));
    } catch (const std::exception &e) {
        g_api_error = e.what();
        return nullptr;
    }
}

cv_canvas *cv_create_tapped(int width, int height, cv_frame_fn on_frame, cv_read_fn on_read,
                            cv_write_fn on_write, void *user)
{
    if (!on_frame) { g_api_error = "cv_create_tapped: on_frame is required"; return nullptr; }
    canvas_ity::canvas *c = new canvas_ity::canvas(width, height, -1, 0, height);
    c->b200()->tap.user = user;
    c->b200()->tap.frame = on_frame;
    c->b200()->tap.read_rgba8 = on_read;
    c->b200()->tap.write_rgba8 = on_write;
    return reinterpret_cast<cv_canvas *>(c);
}

void cv_destroy(cv_canvas *canvas)
{
    if (!canvas) return;
    canvas_ity::canvas *c = front(canvas);
    if (c->b200()->tap.frame) c->b200()->flush();
    delete c;
}

long cv_run_script(cv_canvas *canvas, const uint8_t *script, size_t bytes, uint32_t *queries,
                   int query_capacity, int *n_queries)
{
    if (!canvas || (!script && bytes)) return -1;
    std::vector<cb200_script::query_result> q;
    long n = cb200_script::run_script<canvas_ity::canvas, ns_tag>(*front(canvas), script, bytes, &q);
    if (n_queries) *n_queries = int(q.size());
    for (int i = 0; queries && i < query_capacity && i < int(q.size()); ++i) {
        queries[i * 4 + 0] = q[size_t(i)].code;
        queries[i * 4 + 1] = q[size_t(i)].got_bits;
        queries[i * 4 + 2] = q[size_t(i)].recorded_bits;
        queries[i * 4 + 3] = 0;
    }
    return n;
}

int cv_get_image_data(cv_canvas *canvas, uint8_t *image, int width, int height, int stride,
                      int x, int y)
{
    if (!canvas) return CB200_ERR_BAD_ARG;
    front(canvas)->get_image_data(image, width, height, stride, x, y);
    return CB200_OK;
}

int cv_put_image_data(cv_canvas *canvas, const uint8_t *image, int width, int height,
                      int stride, int x, int y)
{
    if (!canvas) return CB200_ERR_BAD_ARG;
    front(canvas)->put_image_data(image, width, height, stride, x, y);
    return CB200_OK;
}

int cv_is_point_in_path(cv_canvas *canvas, float x, float y)
{
    return canvas && front(canvas)->is_point_in_path(x, y) ? 1 : 0;
}

float cv_measure_text(cv_canvas *canvas, const char *text)
{
    return canvas ? front(canvas)->measure_text(text) : 0.0f;
}

int cv_flush(cv_canvas *canvas)
{
    if (!canvas) return CB200_ERR_BAD_ARG;
    front(canvas)->b200()->flush();
    return CB200_OK;
}

int cv_read_f32(cv_canvas *canvas, float *dst)
{
    if (!canvas || !dst) return CB200_ERR_BAD_ARG;
    canvas_ity::canvas::host_state *s = front(canvas)->b200();
    s->flush();
    if (!s->device) { g_api_error = "cv_read_f32: tapped canvas has no device"; return CB200_ERR_NO_DEVICE; }
    return cb200_read_f32(s->device, dst);
}

cb200_canvas *cv_device(cv_canvas *canvas) { return canvas ? front(canvas)->b200()->device : nullptr; }

const char *cv_last_error(void)
{
    return g_api_error.empty() ? cb200_last_error() : g_api_error.c_str();
}

const char *cv_backend_name(void) { return "b200"; }

}  // extern "C"
