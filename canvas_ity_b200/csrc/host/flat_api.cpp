// flat_api.cpp -- extern "C" entry points of include/canvas_b200_api.h over the
// recording front end (canvas_front.cpp).
#include "front_state.hpp"
#include "script.hpp"

#include "../../../include/canvas_b200_api.h"

#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>
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

// ---- batches ---------------------------------------------------------------------
namespace {
struct owned_frame {                   // deep copy of a lowered frame (the tap's pointers die with the call)
    std::vector<cb200_draw> draws; std::vector<cb200_subpath> subpaths; std::vector<float> points;
    std::vector<cb200_brush> brushes; std::vector<float> colors, stops, dashes;
    std::vector<cb200_image> images; std::vector<uint8_t> texels;
    std::vector<cb200_glyph_atlas> atlases; std::vector<cb200_glyph_inst> glyphs;   // atlas arrays live for the process
    cb200_frame view;
    explicit owned_frame(const cb200_frame &f)
        : draws(f.draws, f.draws + f.n_draws), subpaths(f.subpaths, f.subpaths + f.n_subpaths),
          points(f.points, f.points + 2 * size_t(f.n_points)), brushes(f.brushes, f.brushes + f.n_brushes),
          colors(f.colors, f.colors + 4 * size_t(f.n_colors)), stops(f.stops, f.stops + f.n_colors),
          dashes(f.dashes, f.dashes + f.n_dashes), images(f.images, f.images + f.n_images),
          texels(f.texels, f.texels + f.texel_bytes), atlases(f.atlases, f.atlases + f.n_atlases),
          glyphs(f.glyphs, f.glyphs + f.n_glyphs)
    {
        view = f;
    }
    const cb200_frame *frame()
    {
        view.draws = draws.data(); view.subpaths = subpaths.data(); view.points = points.data();
        view.brushes = brushes.data(); view.colors = colors.data(); view.stops = stops.data();
        view.dashes = dashes.data(); view.images = images.data(); view.texels = texels.data();
        view.atlases = atlases.data(); view.glyphs = glyphs.data();
        return &view;
    }
};
}

struct cv_batch {
    cb200_canvas *device = nullptr;
    int n = 0, width = 0, height = 0;
    std::vector<canvas_ity::canvas *> members;
    std::vector<std::vector<owned_frame *> > pending;      // per canvas, in flush order
    struct slot { cv_batch *batch; int index; };
    std::vector<slot> slots;
};

namespace {
void batch_on_frame(void *user, const cb200_frame *frame)
{
    cv_batch::slot *s = static_cast<cv_batch::slot *>(user);
    s->batch->pending[size_t(s->index)].push_back(new owned_frame(*frame));
}
void batch_on_read(void *user, uint8_t *dst, int w, int h, int stride, int x, int y)
{
    cv_batch::slot *s = static_cast<cv_batch::slot *>(user);
    cv_batch_get_image_data(s->batch, s->index, dst, w, h, stride, x, y);
}
// put_image_data on a member (hpp:3383-3408): ordered after everything drawn so far on any member
void batch_on_write(void *user, const uint8_t *src, int w, int h, int stride, int x, int y)
{
    cv_batch::slot *s = static_cast<cv_batch::slot *>(user);
    int rc = cv_batch_flush(s->batch);
    if (rc == CB200_OK) rc = cb200_batch_write_rgba8(s->batch->device, uint32_t(s->index), src, w, h, stride, x, y);
    if (rc != CB200_OK) {
        // the canvas API has no error channel: a device failure is fatal rather than silently wrong pixels
        fprintf(stderr, "canvas_b200: put_image_data on batch canvas %d failed (%d): %s\n", s->index, rc, cb200_last_error());
        abort();
    }
}
}

extern "C" {

cv_batch *cv_batch_create(int n_canvases, int width, int height, int device)
{
    cv_batch *b = new cv_batch;
    if (cb200_batch_create(n_canvases, width, height, device, &b->device) != CB200_OK) {
        g_api_error = cb200_last_error();
        delete b;
        return nullptr;
    }
    b->n = n_canvases; b->width = width; b->height = height;
    b->pending.resize(size_t(n_canvases));
    b->slots.resize(size_t(n_canvases));
    for (int i = 0; i < n_canvases; ++i) {
        b->slots[size_t(i)].batch = b;
        b->slots[size_t(i)].index = i;
        canvas_ity::canvas *c = new canvas_ity::canvas(width, height, -1, 0, height);
        c->b200()->tap.user = &b->slots[size_t(i)];
        c->b200()->tap.frame = batch_on_frame;
        c->b200()->tap.read_rgba8 = batch_on_read;
        c->b200()->tap.write_rgba8 = batch_on_write;
        b->members.push_back(c);
    }
    return b;
}

cv_canvas *cv_batch_canvas(cv_batch *b, int index)
{
    return b && index >= 0 && index < b->n ? reinterpret_cast<cv_canvas *>(b->members[size_t(index)]) : nullptr;
}

int cv_batch_flush(cv_batch *b)
{
    if (!b) return CB200_ERR_BAD_ARG;
    for (int i = 0; i < b->n; ++i) b->members[size_t(i)]->b200()->flush();
    // round k submits the k-th queued frame of every canvas (almost always there is just one round)
    for (size_t round = 0;; ++round) {
        std::vector<const cb200_frame *> frames;
        std::vector<uint32_t> index;
        for (int i = 0; i < b->n; ++i)
            if (round < b->pending[size_t(i)].size()) {
                frames.push_back(b->pending[size_t(i)][round]->frame());
                index.push_back(uint32_t(i));
            }
        if (frames.empty()) break;
        int rc = cb200_batch_submit(b->device, frames.data(), index.data(), uint32_t(frames.size()));
        if (rc != CB200_OK) { g_api_error = cb200_last_error(); return rc; }
    }
    int rc = cb200_sync(b->device);                        // the frames are about to be freed
    bool submitted = false;
    for (int i = 0; i < b->n; ++i) {
        submitted = submitted || !b->pending[size_t(i)].empty();
        for (owned_frame *f : b->pending[size_t(i)]) delete f;
        b->pending[size_t(i)].clear();
    }
    // clip() allocates a device plane per call: free the ones no member can reach any more (current mask +
    // save stack of every member), or a batch that clips every frame grows without bound
    if (rc == CB200_OK && submitted) {
        std::vector<uint32_t> canvas, slot;
        bool any_clip = false;
        for (int i = 0; i < b->n; ++i) {
            canvas_ity::canvas::host_state *st = b->members[size_t(i)]->b200();
            any_clip = any_clip || st->next_mask > 1;
            if (st->mask) { canvas.push_back(uint32_t(i)); slot.push_back(st->mask); }
            for (size_t k = 0; k < st->saves.size(); ++k)
                if (st->saves[k].mask) { canvas.push_back(uint32_t(i)); slot.push_back(st->saves[k].mask); }
        }
        if (any_clip) rc = cb200_batch_masks_keep(b->device, canvas.data(), slot.data(), uint32_t(canvas.size()));
    }
    return rc;
}

int cv_batch_get_image_data(cv_batch *b, int index, uint8_t *image, int width, int height, int stride, int x, int y)
{
    if (!b || index < 0 || index >= b->n || !image) return CB200_ERR_BAD_ARG;
    int rc = cv_batch_flush(b);
    if (rc != CB200_OK) return rc;
    return cb200_batch_read_rgba8(b->device, uint32_t(index), image, width, height, stride, x, y);
}

int cv_batch_read_f32(cv_batch *b, int index, float *dst)
{
    if (!b || index < 0 || index >= b->n || !dst) return CB200_ERR_BAD_ARG;
    int rc = cv_batch_flush(b);
    if (rc != CB200_OK) return rc;
    return cb200_batch_read_f32(b->device, uint32_t(index), dst);
}

cb200_canvas *cv_batch_device(cv_batch *b) { return b ? b->device : nullptr; }

void cv_batch_destroy(cv_batch *b)
{
    if (!b) return;
    for (canvas_ity::canvas *c : b->members) delete c;      // tapped members drop unflushed draws
    for (auto &q : b->pending) for (owned_frame *f : q) delete f;
    cb200_canvas_destroy(b->device);
    delete b;
}

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

int cv_points_in_path(cv_canvas *canvas, const float *xy, int n, uint8_t *inside)
{
    if (!canvas || n < 0 || (n && (!xy || !inside))) return CB200_ERR_BAD_ARG;
    canvas_ity::canvas::host_state *s = front(canvas)->b200();
    if (!s->device) { g_api_error = "cv_points_in_path: tapped canvas has no device"; return CB200_ERR_NO_DEVICE; }
    std::vector<float> edges;
    canvas_ity::flattened_path_edges(s, edges);
    int rc = cb200_hit_test(s->device, edges.data(), uint32_t(edges.size() / 4), xy, uint32_t(n), inside, nullptr);
    if (rc != CB200_OK) g_api_error = cb200_last_error();
    return rc;
}

long cv_path_edges(cv_canvas *canvas, float *edges, long capacity)
{
    if (!canvas) return -1;
    std::vector<float> all;
    canvas_ity::flattened_path_edges(front(canvas)->b200(), all);
    long n = long(all.size() / 4);
    if (edges && capacity > 0) memcpy(edges, all.data(), sizeof(float) * 4 * size_t(std::min(n, capacity)));
    return n;
}

int cv_write_tga(cv_canvas *canvas, const char *path)
{
    if (!canvas || !path) return CB200_ERR_BAD_ARG;
    canvas_ity::canvas::host_state *s = front(canvas)->b200();
    s->flush();
    if (!s->device) { g_api_error = "cv_write_tga: tapped canvas has no device"; return CB200_ERR_NO_DEVICE; }
    const int w = s->width, h = s->height;
    if (w > 0xffff || h > 0xffff) { g_api_error = "cv_write_tga: TGA holds at most 65535 x 65535 pixels"; return CB200_ERR_BAD_ARG; }
    const size_t bytes = size_t(w) * size_t(h) * 4;
    uint8_t *pixels = static_cast<uint8_t *>(cb200_host_alloc(bytes));        // page-locked: one DMA, no bounce
    if (!pixels) { g_api_error = "cv_write_tga: out of host memory"; return CB200_ERR_OOM; }
    int rc = cb200_read_bgra8(s->device, pixels, w, h, 4 * w, 0, 0);
    if (rc == CB200_OK) {
        const unsigned char header[18] = { 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                           (unsigned char)(w & 255), (unsigned char)(w >> 8),
                                           (unsigned char)(h & 255), (unsigned char)(h >> 8), 32, 40 };
        FILE *f = fopen(path, "wb");
        if (!f || fwrite(header, 1, sizeof header, f) != sizeof header || fwrite(pixels, 1, bytes, f) != bytes) {
            g_api_error = std::string("cv_write_tga: cannot write ") + path;
            rc = CB200_ERR_BAD_ARG;
        }
        if (f) fclose(f);
    } else g_api_error = cb200_last_error();
    cb200_host_free(pixels);
    return rc;
}

int cv_write_png(cv_canvas *canvas, const char *path)
{
    if (!canvas || !path) return CB200_ERR_BAD_ARG;
    canvas_ity::canvas::host_state *s = front(canvas)->b200();
    s->flush();
    if (!s->device) { g_api_error = "cv_write_png: tapped canvas has no device"; return CB200_ERR_NO_DEVICE; }
    size_t bytes = 0;
    int rc = cb200_encode_png(s->device, nullptr, 0, &bytes);
    if (rc != CB200_OK) { g_api_error = cb200_last_error(); return rc; }
    uint8_t *file = static_cast<uint8_t *>(cb200_host_alloc(bytes));          // page-locked: one DMA, no bounce
    if (!file) { g_api_error = "cv_write_png: out of host memory"; return CB200_ERR_OOM; }
    rc = cb200_encode_png(s->device, file, bytes, nullptr);
    if (rc == CB200_OK) {
        FILE *f = fopen(path, "wb");
        if (!f || fwrite(file, 1, bytes, f) != bytes) {
            g_api_error = std::string("cv_write_png: cannot write ") + path;
            rc = CB200_ERR_BAD_ARG;
        }
        if (f) fclose(f);
    } else g_api_error = cb200_last_error();
    cb200_host_free(file);
    return rc;
}

int cv_set_text_instancing(cv_canvas *canvas, int on)
{
    if (!canvas) return CB200_ERR_BAD_ARG;
    front(canvas)->b200()->instanced_text = on != 0;
    return CB200_OK;
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
