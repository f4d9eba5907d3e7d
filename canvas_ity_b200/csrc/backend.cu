// backend.cu -- the C ABI of include/canvas_b200.h: device canvases, frame
// translation/upload and the fixed per-frame launch sequence.
//
// A submitted cb200_frame is translated on the host into device records
// (draw_rec, subpath_rec, flatten units, raster jobs, ...), packed into ONE pinned
// staging blob, copied with ONE cudaMemcpyAsync and processed by a fixed sequence
// of frame-level kernels on the canvas' stream (see device/common.cuh).  Every
// data-dependent size stays on the device; capacities are checked by the kernels
// themselves (frame_header::overflow) and, if one was too small, the frame is
// re-run with larger buffers before its result is ever observed -- the
// compositor does not touch the framebuffer of an overflowed frame.
//
// There is no CPU fallback here: without a CUDA device cb200_canvas_create fails.
#include "device/frame.cuh"
#include "device/edge_clip.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace cb200;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &what)
{
    g_error = what;
    return code;
}

#define CK(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? CB200_ERR_OOM : CB200_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));              \
    } while (0)

template <class T>
struct dev_buf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host-side image of everything uploaded for one frame.
struct staged_frame {
    std::vector<draw_rec> draws;
    std::vector<subpath_rec> subpaths;
    std::vector<unit_rec> units;
    std::vector<float> points;
    std::vector<brush_rec> brushes;
    std::vector<float> colors, stops, dashes;
    std::vector<dash_item> dash_items;
    std::vector<uint2> draw_src;
    std::vector<stroke_src> sources;
    std::vector<job_rec> jobs;
    std::vector<uint32_t> shadow_jobs;
    std::vector<uint8_t> texels_u8;
    std::vector<float *> mask_table;
    std::vector<uint2> canvas_jobs;              // per canvas: first job, job count
    // glyph instances: expanded on the device into the point pool right after the uploaded points
    std::vector<glyph_inst_rec> glyph_insts;
    std::vector<cb200_glyph_atlas> atlas_views;  // distinct atlases (by id) the instances refer to
    std::vector<atlas_dev> atlas_table;          // device arrays of each, filled by upload_frame
    uint32_t n_glyph_points = 0;
    uint64_t n_texels = 0;
    int key_bits = 0, bits_x = 0, bits_y = 0;
    uint32_t n_dash_subpath_cap = 0;
    bool valid = false;
};

int bits_for(uint32_t max_value)
{
    int b = 1;
    while ((uint64_t(1) << b) <= max_value) ++b;
    return b;
}

}  // namespace

struct cb200_canvas {
    int device = 0, width = 0, height = 0, band_y0 = 0, band_rows = 0;
    int n_canvases = 1, slot_rows = 0;        // batch: canvases stacked vertically, slot_rows (multiple of 32) apart
    std::map<uint32_t, std::map<uint32_t, uint32_t> > batch_masks;   // canvas -> local clip slot -> batch slot
    uint32_t next_batch_mask = 1;
    cudaStream_t stream = nullptr;
    float4 *fb = nullptr;
    std::map<uint32_t, float *> masks;

    // input blob (device + pinned host mirror)
    dev_buf<uint8_t> blob;
    uint8_t *pinned = nullptr;
    size_t pinned_cap = 0;
    frame_header *pinned_hdr = nullptr;       // readback of the header after a frame

    // device work buffers
    dev_buf<uint32_t> unit_count, unit_offset, pt_loop, dash_pts_count, dash_sub_count, dash_tail,
        half_count, half_offset, half_unit_off, half_dirty, stroke_unit_pts, half_last, visit_prev, long_rows, piece_job, piece_rows, piece_rlo, piece_row_off, row_runs, row_piece, te_flags, te_job,
        te_first, te_mask, partials, sort_hist;
    dev_buf<float2> pts;
    dev_buf<loop_span> loops;
    dev_buf<stroke_src> sources;
    dev_buf<float4> pieces, texels;
    dev_buf<comp_rec> comp;
    dev_buf<uint2> job_box;
    dev_buf<uint32_t> job_te, blur_units, row_jobs, row_job_count, loop_mark, job_run_begin;
    dev_buf<uint2> box_loops;
    dev_buf<leak_rec> leaks;  uint32_t cap_leaks = 0;
    dev_buf<uint32_t> tile_cover;
    dev_buf<uint64_t> keys0, keys1;
    dev_buf<float> vals0, vals1, cumulative, te_backdrop, planes, planes_tmp;
    dev_buf<uint8_t> rgba8, visit_close;
    uint8_t *pinned_rgba8 = nullptr;
    size_t pinned_rgba8_cap = 0;

    uint32_t cap_pts = 0, cap_items = 0, cap_rows = 0, cap_runs = 0, cap_tiles = 0, cap_sources = 0,
             cap_dash_subpaths = 0;
    uint64_t cap_planes = 0;

    // PNG encode (cb200_encode_png): per-size constant tables, the file image, the checksum accumulators
    dev_buf<uint32_t> png_tables, png_row_crc;  dev_buf<uint8_t> png_out;  dev_buf<png_sums> png_acc;  bool png_tables_ready = false;
    // bulk hit testing (cb200_hit_test)
    dev_buf<float4> hit_edges;  dev_buf<float2> hit_queries;  dev_buf<int2> hit_acc;  dev_buf<uint8_t> hit_inside;
    // device mirrors of the glyph atlases seen so far (append-only, keyed by atlas id)
    struct atlas_mirror {
        dev_buf<cb200_glyph_outline> outlines; dev_buf<cb200_glyph_seg> segs; dev_buf<float2> points;
        uint32_t n_outlines = 0, n_segs = 0, n_points = 0;
    };
    std::map<uint64_t, atlas_mirror> atlases;

    staged_frame staged;
    device_frame df;
    canvas_target target;
    bool pending = false;                     // a frame was launched and not yet verified
    bool resident = false;                    // staged frame came from cb200_frame_upload
    size_t hdr_offset = 0, hdr_pristine_offset = 0;
    cudaEvent_t ev[10];
    uint32_t row_stride = 0;  bool use_row_lists = false;
    bool stage_timing = true;          // cb200_set_stage_timing
    bool read_bgra = false;            // channel order of the readback in progress (cb200_read_bgra8)
    bool clear_pending = false;        // cb200_clear / replay(clear): folded into the next frame, or applied by settle()
    bool inflight_clear = false;       // the frame in flight starts from a cleared canvas (kept for overflow re-runs)
    bool replay_verified = false;      // the resident frame has completed once with the current capacities
    // replays of a verified resident frame run as ONE cudaGraphLaunch instead of ~45 stream calls
    // (the host enqueue cost, not the GPU, bounds frames/s when several canvases replay concurrently);
    // [0] without, [1] with the folded-in clear.  Dropped whenever the frame's buffers change.
    cudaGraphExec_t replay_graph[2] = {nullptr, nullptr};
    bool graph_replay = true;          // cb200_set_graph_replay
    bool graph_broken = false;         // capture or instantiation failed once: keep to stream launches
    uint32_t graph_replays = 0;
    // composite kernel times of the most recent frames (bench: roofline over the timed region)
    static const int kCompRing = 256;
    cudaEvent_t comp_ev[2 * kCompRing] = {};
    bool comp_ev_valid[kCompRing] = {};   // false for frames that ran inside a graph (no per-frame events)
    uint64_t frames_run = 0, timer_frame0 = 0;
    cudaEvent_t timer_ev[2] = {};
    bool timer_stopped = false;       // cb200_timer_stop already recorded the end event
    std::vector<cudaEvent_t> chunk_events;
    cb200_stats stats;
    uint64_t launches = 0;
};

namespace {

size_t fb_rows(const cb200_canvas *cv) { return cv->n_canvases > 1 ? size_t(cv->n_canvases) * size_t(cv->slot_rows) : size_t(cv->band_rows); }

// ---------------------------------------------------------------- staging ----

// Extended-box blur parameters of one draw, hpp:2402-2405, 2453-2459.
void blur_params(float blur, int &radius, int &border, float &w1, float &w2)
{
    float sigma2 = 0.25f * blur * blur;
    size_t r = static_cast<size_t>(0.5f * sqrtf(4.0f * sigma2 + 1.0f) - 0.5f);
    radius = int(r);
    border = 3 * (radius + 1);
    float alpha = static_cast<float>(2 * r + 1) * (static_cast<float>(r * (r + 1)) - sigma2) /
                  (2.0f * sigma2 - static_cast<float>(6 * (r + 1) * (r + 1)));
    float div = 2.0f * (alpha + static_cast<float>(r)) + 1.0f;
    w1 = alpha / div;
    w2 = (1.0f - alpha) / div;
}

affine to_affine(const float *m) { affine a = { m[0], m[1], m[2], m[3], m[4], m[5] }; return a; }

// Translate one or more lowered frames (one per canvas of a batch, canvas indices ascending)
// into the device records of ONE device frame.  Indices of later frames are rebased onto the
// shared pools; clip-mask slots are renamed to batch-wide slots.
int stage_frames(cb200_canvas *cv, const cb200_frame *const *frames, const uint32_t *canvas_index,
                 uint32_t n_frames, staged_frame &sf)
{
    sf = staged_frame();
    sf.canvas_jobs.assign(size_t(cv->n_canvases), make_uint2(0, 0));
    int max_pad = 0;
    uint32_t last_canvas = 0;
    for (uint32_t fi = 0; fi < n_frames; ++fi) {
        const cb200_frame *in = frames[fi];
        const uint32_t canvas = canvas_index ? canvas_index[fi] : 0;
        if (!in) return fail(CB200_ERR_BAD_ARG, "null frame");
        if (canvas >= uint32_t(cv->n_canvases)) return fail(CB200_ERR_BAD_ARG, "canvas index out of range");
        if (fi && canvas <= last_canvas) return fail(CB200_ERR_BAD_ARG, "batch frames must have ascending canvas indices");
        last_canvas = canvas;
        if (in->n_draws && !in->draws) return fail(CB200_ERR_BAD_ARG, "frame.draws is null");
        const uint32_t draw_base = uint32_t(sf.draws.size()), sub_base = uint32_t(sf.subpaths.size());
        const uint32_t pt_base = uint32_t(sf.points.size() / 2), brush_base = uint32_t(sf.brushes.size());
        const uint32_t color_base = uint32_t(sf.stops.size()), dash_base = uint32_t(sf.dashes.size());
        const uint32_t job_base = uint32_t(sf.jobs.size());
        sf.draws.resize(draw_base + in->n_draws);
        sf.subpaths.resize(sub_base + in->n_subpaths);
        sf.draw_src.resize(draw_base + in->n_draws, make_uint2(0, 0));
        // glyph instances: rebase onto the shared instance region and the batch-wide atlas list
        const uint32_t glyph_pt_base = sf.n_glyph_points;
        if ((in->n_glyphs && !in->glyphs) || (in->n_atlases && !in->atlases) || (in->n_glyphs && !in->n_atlases))
            return fail(CB200_ERR_BAD_ARG, "frame.glyphs / atlases is null");
        std::vector<uint32_t> atlas_slot(in->n_atlases);
        for (uint32_t a = 0; a < in->n_atlases; ++a) {
            uint32_t slot = 0;
            while (slot < sf.atlas_views.size() && sf.atlas_views[slot].id != in->atlases[a].id) ++slot;
            if (slot == sf.atlas_views.size()) sf.atlas_views.push_back(in->atlases[a]);
            else if (in->atlases[a].n_outlines > sf.atlas_views[slot].n_outlines) sf.atlas_views[slot] = in->atlases[a];   // the later snapshot
            atlas_slot[a] = slot;
        }
        for (uint32_t g = 0; g < in->n_glyphs; ++g) {
            const cb200_glyph_inst &gi = in->glyphs[g];
            if (gi.atlas >= in->n_atlases || gi.outline >= in->atlases[gi.atlas].n_outlines)
                return fail(CB200_ERR_BAD_ARG, "glyph instance refers to an unknown outline");
            const cb200_glyph_outline &o = in->atlases[gi.atlas].outlines[gi.outline];
            if (uint64_t(gi.first_point) + o.out_points > in->n_glyph_points)
                return fail(CB200_ERR_BAD_ARG, "glyph instance points out of range");
            glyph_inst_rec r = { atlas_slot[gi.atlas], gi.outline, glyph_pt_base + gi.first_point, 0, to_affine(gi.m) };
            sf.glyph_insts.push_back(r);
        }
        sf.n_glyph_points += in->n_glyph_points;
        for (uint32_t s = 0; s < in->n_subpaths; ++s) {
            const cb200_subpath &sp = in->subpaths[s];
            if (uint64_t(sp.first_point) + 1 + 3ull * sp.n_cubics > (sp.instanced ? in->n_glyph_points : in->n_points))
                return fail(CB200_ERR_BAD_ARG, "subpath points out of range");
            // instanced subpaths: bit 31 marks "relative to the instance region" until every frame's
            // uploaded points are counted (fixed up below)
            subpath_rec r = { sp.instanced ? (0x80000000u | (glyph_pt_base + sp.first_point)) : pt_base + sp.first_point,
                              sp.n_cubics, sp.closed, 0xffffffffu, 0 };
            sf.subpaths[sub_base + s] = r;
        }
        std::map<uint32_t, uint32_t> *slots = cv->n_canvases > 1 ? &cv->batch_masks[canvas] : nullptr;
        auto global_slot = [&](uint32_t local, bool define) -> uint32_t {
            if (!slots || local == 0) return local;
            if (define) return (*slots)[local] = cv->next_batch_mask++;
            std::map<uint32_t, uint32_t>::iterator it = slots->find(local);
            return it == slots->end() ? 0xffffffffu : it->second;
        };
        for (uint32_t i = 0; i < in->n_draws; ++i) {
            const cb200_draw &d = in->draws[i];
            const uint32_t di = draw_base + i;
            if (uint64_t(d.first_subpath) + d.n_subpaths > in->n_subpaths)
                return fail(CB200_ERR_BAD_ARG, "draw subpaths out of range");
            if (d.kind != CB200_CLIP && d.brush >= in->n_brushes) return fail(CB200_ERR_BAD_ARG, "draw brush out of range");
            if (uint64_t(d.first_dash) + d.n_dash > in->n_dashes) return fail(CB200_ERR_BAD_ARG, "draw dashes out of range");
            draw_rec r;
            memset(&r, 0, sizeof r);
            r.kind = d.kind; r.op = d.op;
            r.first_subpath = sub_base + d.first_subpath; r.n_subpaths = d.n_subpaths;
            r.brush = brush_base + d.brush;
            r.mask_src = global_slot(d.mask_src, false);
            if (r.mask_src == 0xffffffffu) return fail(CB200_ERR_BAD_ARG, "draw reads a clip-mask slot that was never written");
            r.mask_dst = d.kind == CB200_CLIP ? global_slot(d.mask_dst, true) : 0;
            r.cap = d.cap; r.join = d.join;
            r.first_dash = dash_base + d.first_dash; r.n_dash = d.kind == CB200_STROKE ? d.n_dash : 0;
            r.dash_offset = d.dash_offset; r.global_alpha = d.global_alpha;
            r.line_width = d.line_width; r.miter_limit = d.miter_limit;
            r.forward = to_affine(d.forward); r.inverse = to_affine(d.inverse);
            memcpy(r.shadow_color, d.shadow_color, sizeof r.shadow_color);
            r.shadow_dx = d.shadow_offset_x; r.shadow_dy = d.shadow_offset_y; r.shadow_blur = d.shadow_blur;
            r.angular = d.kind == CB200_STROKE ? stroke_angular(d.line_width) : -1.0f;
            r.canvas = canvas;
            r.first_unit = uint32_t(sf.units.size());
            for (uint32_t s = 0; s < d.n_subpaths; ++s) {
                subpath_rec &sp = sf.subpaths[sub_base + d.first_subpath + s];
                if (sp.draw != 0xffffffffu) return fail(CB200_ERR_BAD_ARG, "subpath shared by two draws");
                sp.draw = di;
                sp.first_unit = uint32_t(sf.units.size());
                for (uint32_t k = 0; k <= sp.n_cubics; ++k) {
                    unit_rec u = { sub_base + d.first_subpath + s, k };
                    sf.units.push_back(u);
                }
            }
            r.n_units = uint32_t(sf.units.size()) - r.first_unit;
            // stroke sources: un-dashed strokes list their subpaths now, dashed ones are
            // appended on the device by K2
            if (d.kind == CB200_STROKE && r.n_dash)
                for (uint32_t s = 0; s < d.n_subpaths; ++s) {
                    dash_item it = { sub_base + d.first_subpath + s, (s == 0 ? 1u : 0u) | (s + 1 == d.n_subpaths ? 2u : 0u) };
                    sf.dash_items.push_back(it);
                }
            sf.draws[di] = r;
            // jobs in composite order: shadow (if any) before the draw itself
            bool shadow = d.kind != CB200_CLIP && d.shadow_color[3] != 0.0f &&
                          (d.shadow_blur != 0.0f || d.shadow_offset_x != 0.0f || d.shadow_offset_y != 0.0f);
            if (shadow) {
                job_rec j;
                memset(&j, 0, sizeof j);
                j.draw = di; j.kind = JOB_SHADOW; j.canvas = canvas;
                blur_params(d.shadow_blur, j.radius, j.border, j.w1, j.w2);
                j.off_x = static_cast<float>(j.border) + d.shadow_offset_x;
                j.off_y = static_cast<float>(j.border) + d.shadow_offset_y;
                j.pad = 2 * j.border;
                max_pad = std::max(max_pad, j.pad);
                sf.shadow_jobs.push_back(uint32_t(sf.jobs.size()));
                sf.jobs.push_back(j);
            }
            job_rec j;
            memset(&j, 0, sizeof j);
            j.draw = di; j.canvas = canvas;
            j.kind = d.kind == CB200_CLIP ? JOB_CLIP : JOB_MAIN;
            // occlusion-culling candidate: where its coverage is exactly 1 this draw REPLACES the pixel
            // (hpp:2583-2591 with cov = vis = 1): solid colour, unclipped, and either source_copy or
            // source_over with global_alpha == colour alpha == 1
            if (j.kind == JOB_MAIN && d.mask_src == 0 && in->brushes[d.brush].type == CB200_BRUSH_COLOR &&
                in->brushes[d.brush].n_colors == 1 && in->brushes[d.brush].first_color < in->n_colors) {
                float a = in->colors[4 * size_t(in->brushes[d.brush].first_color) + 3];
                if (d.op == 2u || (d.op == 14u && d.global_alpha == 1.0f && a == 1.0f)) j.opaque = 1;
            }
            sf.jobs.push_back(j);
        }
        sf.canvas_jobs[canvas] = make_uint2(job_base, uint32_t(sf.jobs.size()) - job_base);
        // static stroke sources, grouped by draw (keeps every draw's K3 output contiguous)
        for (uint32_t i = 0; i < in->n_draws; ++i) {
            const cb200_draw &d = in->draws[i];
            const uint32_t di = draw_base + i;
            if (d.kind != CB200_STROKE || sf.draws[di].n_dash) continue;
            sf.draw_src[di].x = uint32_t(sf.sources.size());
            for (uint32_t s = 0; s < d.n_subpaths; ++s) {
                stroke_src src = { sub_base + d.first_subpath + s,
                                   di | (in->subpaths[d.first_subpath + s].closed ? 0x80000000u : 0u) };
                sf.sources.push_back(src);
            }
            sf.draw_src[di].y = uint32_t(sf.sources.size());
        }
        if (in->n_points) sf.points.insert(sf.points.end(), in->points, in->points + 2 * size_t(in->n_points));
        if (in->n_colors) {
            sf.colors.insert(sf.colors.end(), in->colors, in->colors + 4 * size_t(in->n_colors));
            sf.stops.insert(sf.stops.end(), in->stops, in->stops + in->n_colors);
        }
        if (in->n_dashes) sf.dashes.insert(sf.dashes.end(), in->dashes, in->dashes + in->n_dashes);
        // brushes; pattern images are converted to float4 texels on the device
        std::vector<uint64_t> image_texel_base(in->n_images);
        for (uint32_t k = 0; k < in->n_images; ++k) {
            const cb200_image &im = in->images[k];
            if (im.width <= 0 || im.height <= 0 ||
                im.texel_offset + 4ull * uint64_t(im.width) * uint64_t(im.height) > in->texel_bytes)
                return fail(CB200_ERR_BAD_ARG, "image out of range");
            image_texel_base[k] = sf.n_texels;
            size_t bytes = 4 * size_t(im.width) * size_t(im.height);
            sf.texels_u8.insert(sf.texels_u8.end(), in->texels + im.texel_offset, in->texels + im.texel_offset + bytes);
            sf.n_texels += uint64_t(im.width) * uint64_t(im.height);
        }
        sf.brushes.resize(brush_base + in->n_brushes);
        for (uint32_t k = 0; k < in->n_brushes; ++k) {
            const cb200_brush &b = in->brushes[k];
            brush_rec r;
            memset(&r, 0, sizeof r);
            r.type = b.type; r.flags = b.flags; r.first_color = color_base + b.first_color; r.n_colors = b.n_colors;
            r.sx = b.start[0]; r.sy = b.start[1]; r.ex = b.end[0]; r.ey = b.end[1];
            r.r0 = b.start_radius; r.r1 = b.end_radius; r.repetition = b.repetition;
            if (b.type == CB200_BRUSH_PATTERN) {
                if (b.image >= in->n_images) return fail(CB200_ERR_BAD_ARG, "brush image out of range");
                r.width = in->images[b.image].width;
                r.height = in->images[b.image].height;
                r.texel_offset = image_texel_base[b.image];
                r.n_colors = 1;
            } else if (uint64_t(b.first_color) + b.n_colors > in->n_colors)
                return fail(CB200_ERR_BAD_ARG, "brush colours out of range");
            sf.brushes[brush_base + k] = r;
        }
    }
    if (cv->n_canvases == 1) sf.canvas_jobs[0] = make_uint2(0, uint32_t(sf.jobs.size()));
    // the instance region follows ALL uploaded points
    const uint32_t uploaded_points = uint32_t(sf.points.size() / 2);
    if (uint64_t(uploaded_points) + sf.n_glyph_points >= 0x80000000ull) return fail(CB200_ERR_BAD_ARG, "too many path points");
    for (subpath_rec &sp : sf.subpaths)
        if (sp.first_point & 0x80000000u) sp.first_point = uploaded_points + (sp.first_point & 0x7fffffffu);
    for (glyph_inst_rec &g : sf.glyph_insts) g.first_point += uploaded_points;
    // sort key layout: job | y | x   (y, x local to the job's canvas)
    sf.bits_x = bits_for(uint32_t(cv->width + max_pad + 1));
    sf.bits_y = bits_for(uint32_t(cv->height + max_pad + 1));
    sf.key_bits = sf.bits_x + sf.bits_y + bits_for(uint32_t(sf.jobs.size()));
    if (sf.key_bits > 64) return fail(CB200_ERR_BAD_ARG, "too many jobs for the sort key");
    sf.valid = true;
    return CB200_OK;
}

int stage_frame(cb200_canvas *cv, const cb200_frame *in, staged_frame &sf)
{
    if (cv->n_canvases != 1) return fail(CB200_ERR_BAD_ARG, "use cb200_batch_submit on a batch");
    return stage_frames(cv, &in, nullptr, 1, sf);
}

// ------------------------------------------------------------ capacities ----

int ensure_capacity(cb200_canvas *cv, const staged_frame &sf, const frame_header *seen)
{
    // first guesses scale with the input; after an overflow the header tells the truth
    uint32_t n_units = uint32_t(sf.units.size());
    uint32_t want_pts = std::max<uint32_t>(1u << 18, n_units * 192u);
    uint32_t want_sources = uint32_t(sf.sources.size()) + std::max<uint32_t>(4096u, uint32_t(sf.dash_items.size()) * 64u);
    uint32_t want_dash_sub = sf.dash_items.empty() ? 0u : std::max<uint32_t>(4096u, uint32_t(sf.dash_items.size()) * 64u);
    uint32_t want_items = want_pts * 2, want_rows = 1u << 22, want_runs = 1u << 23, want_tiles = 1u << 17;
    uint32_t want_leaks = 4096;
    uint64_t want_planes = sf.shadow_jobs.empty() ? 0 : uint64_t(cv->width + 64) * uint64_t(cv->height + 64) * 2;
    if (getenv("CB200_TEST_SMALL_CAPS")) {
        // test hook: start with tiny queues so that every frame exercises the overflow -> regrow -> re-run path
        want_pts = 64; want_items = 64; want_rows = 64; want_runs = 64; want_tiles = 16; want_leaks = 1;
        want_planes = sf.shadow_jobs.empty() ? 0 : 64;
        want_dash_sub = sf.dash_items.empty() ? 0u : 4u;
        want_sources = uint32_t(sf.sources.size()) + 4;
    }
    if (seen) {
        uint32_t total_pts = seen->n_line_points + seen->n_dash_points + seen->n_stroke_points;
        want_pts = std::max(want_pts, total_pts + total_pts / 4 + 1024);
        want_sources = std::max(want_sources, seen->n_sources + seen->n_sources / 4 + 64);
        want_dash_sub = std::max(want_dash_sub, seen->n_dash_subpaths + seen->n_dash_subpaths / 4 + 64);
        want_items = std::max(want_items, seen->n_items + seen->n_items / 4 + 1024);
        want_rows = std::max(want_rows, seen->n_row_items + seen->n_row_items / 4 + 1024);
        want_runs = std::max(want_runs, seen->n_runs + seen->n_runs / 4 + 1024);
        want_tiles = std::max(want_tiles, seen->n_tile_entries + seen->n_tile_entries / 4 + 1024);
        want_planes = std::max<uint64_t>(want_planes, seen->plane_floats + seen->plane_floats / 4 + 1024);
        if (seen->overflow & (OVF_POINTS | OVF_DASH)) want_pts = std::max(want_pts, cv->cap_pts * 2);
        want_leaks = std::max(want_leaks, seen->n_leaks + seen->n_leaks / 4 + 16);
        if (seen->overflow & OVF_DASH) {
            want_sources = std::max(want_sources, cv->cap_sources * 2);
            want_dash_sub = std::max(want_dash_sub, cv->cap_dash_subpaths * 2);
        }
    }
    want_pts = std::max(want_pts, cv->cap_pts);
    want_sources = std::max(want_sources, cv->cap_sources);
    want_dash_sub = std::max(want_dash_sub, cv->cap_dash_subpaths);
    want_items = std::max(want_items, std::max(cv->cap_items, want_pts * 2));
    want_rows = std::max(want_rows, cv->cap_rows);
    want_runs = std::max(want_runs, cv->cap_runs);
    want_tiles = std::max(want_tiles, cv->cap_tiles);
    want_planes = std::max(want_planes, cv->cap_planes);
    want_leaks = std::max(want_leaks, cv->cap_leaks);

    CK(cv->unit_count.reserve(n_units + 1));
    CK(cv->unit_offset.reserve(n_units + 2));
    CK(cv->pts.reserve(want_pts));
    CK(cv->pt_loop.reserve(want_pts));
    CK(cv->dash_pts_count.reserve(sf.dash_items.size() + 1));
    CK(cv->dash_sub_count.reserve(sf.dash_items.size() + 1));
    CK(cv->dash_tail.reserve(sf.dash_items.size() + 1));
    CK(cv->sources.reserve(want_sources));
    CK(cv->half_count.reserve(2 * size_t(want_sources) + 2));
    CK(cv->half_offset.reserve(2 * size_t(want_sources) + 4));
    CK(cv->half_unit_off.reserve(2 * size_t(want_sources) + 4));
    CK(cv->half_dirty.reserve(2 * size_t(want_sources) + 4));
    CK(cv->stroke_unit_pts.reserve(size_t(want_pts) + 6 * size_t(want_sources) + 4));
    CK(cv->visit_prev.reserve(size_t(want_pts) + 6 * size_t(want_sources) + 4));
    CK(cv->visit_close.reserve(size_t(want_pts) + 6 * size_t(want_sources) + 4));
    CK(cv->half_last.reserve(2 * size_t(want_sources) + 4));
    CK(cv->loops.reserve(sf.subpaths.size() + want_dash_sub + 2 * size_t(want_sources) + 2));
    if (!sf.shadow_jobs.empty()) {
        CK(cv->loop_mark.reserve(cv->loops.cap));
        CK(cv->box_loops.reserve(cv->loops.cap));
    }
    CK(cv->pieces.reserve(3 * size_t(want_items)));
    CK(cv->piece_job.reserve(3 * size_t(want_items)));
    CK(cv->piece_rows.reserve(3 * size_t(want_items)));
    CK(cv->piece_rlo.reserve(3 * size_t(want_items)));
    CK(cv->piece_row_off.reserve(3 * size_t(want_items)));
    CK(cv->row_runs.reserve(want_rows));
    CK(cv->row_piece.reserve(want_rows));
    CK(cv->keys0.reserve(want_runs));
    CK(cv->keys1.reserve(want_runs));
    CK(cv->vals0.reserve(want_runs));
    CK(cv->vals1.reserve(want_runs));
    CK(cv->cumulative.reserve(want_runs));
    CK(cv->long_rows.reserve(want_runs / 32 + 64));
    CK(cv->leaks.reserve(want_leaks));
    {   // one word per tile of the target (band or stacked batch)
        const size_t tiles_x = size_t((cv->width + kTile - 1) / kTile);
        const size_t rows = cv->n_canvases > 1 ? size_t(cv->n_canvases) * size_t(cv->slot_rows / kTile)
                                                : size_t((cv->band_y0 + cv->band_rows - 1) / kTile - cv->band_y0 / kTile + 1);
        CK(cv->tile_cover.reserve(tiles_x * rows));
    }
    CK(cv->te_flags.reserve(want_tiles));
    CK(cv->te_job.reserve(want_tiles));
    CK(cv->te_backdrop.reserve(size_t(want_tiles) * kTile));
    CK(cv->te_first.reserve(size_t(want_tiles) * kTile));
    CK(cv->te_mask.reserve(size_t(want_tiles) * kTile));
    CK(cv->planes.reserve(want_planes));
    CK(cv->planes_tmp.reserve(want_planes));
    CK(cv->partials.reserve(8 * kGrid));
    CK(cv->comp.reserve(sf.jobs.size() + 1));
    CK(cv->job_box.reserve(sf.jobs.size() + 1));
    CK(cv->job_te.reserve(sf.jobs.size() + 1));
    CK(cv->job_run_begin.reserve(sf.jobs.size() + 2));
    CK(cv->blur_units.reserve(2 * (sf.shadow_jobs.size() + 1) + 2));
    {   // tile-row job lists: rows of the target x the largest job count of a canvas (skipped when huge)
        uint32_t most = 0;
        for (const uint2 &cj : sf.canvas_jobs) most = std::max(most, cj.y);
        const size_t rows = cv->n_canvases > 1 ? size_t(cv->n_canvases) * size_t(cv->slot_rows / kTile)
                                                : size_t((cv->band_y0 + cv->band_rows - 1) / kTile - cv->band_y0 / kTile + 1);
        cv->row_stride = (most + 31u) & ~31u;
        cv->use_row_lists = most > 32 && rows * cv->row_stride <= (size_t(1) << 25);
        if (cv->use_row_lists) {
            CK(cv->row_jobs.reserve(rows * cv->row_stride));
            CK(cv->row_job_count.reserve(rows));
        }
    }
    CK(cv->sort_hist.reserve(512 * kGrid + 512));
    CK(cv->texels.reserve(std::max<uint64_t>(sf.n_texels, 1)));
    cv->cap_pts = want_pts; cv->cap_sources = want_sources; cv->cap_dash_subpaths = want_dash_sub;
    cv->cap_items = want_items; cv->cap_rows = want_rows; cv->cap_runs = want_runs;
    cv->cap_tiles = want_tiles; cv->cap_planes = want_planes; cv->cap_leaks = want_leaks;
    return CB200_OK;
}

// Pack the staged frame into the pinned blob, upload it and point device_frame
// at the pieces.  Layout: [header][pristine header][arrays...], 256 B aligned.
template <class T>
size_t place(std::vector<std::pair<size_t, std::pair<const void *, size_t> > > &plan, size_t &at,
             const std::vector<T> &v)
{
    size_t off = at;
    plan.push_back(std::make_pair(off, std::make_pair(static_cast<const void *>(v.data()), v.size() * sizeof(T))));
    at = align_up(at + std::max<size_t>(v.size() * sizeof(T), 16), 256);
    return off;
}

void drop_replay_graphs(cb200_canvas *cv);

int upload_frame(cb200_canvas *cv)
{
    drop_replay_graphs(cv);                               // captured pointers and sizes are about to change
    staged_frame &sf = cv->staged;
    // clip masks written by this frame need planes before the table is built
    uint32_t max_slot = 0;
    for (const draw_rec &d : sf.draws) {
        max_slot = std::max(max_slot, std::max(d.mask_src, d.mask_dst));
        if (d.kind == CB200_CLIP && !cv->masks.count(d.mask_dst)) {
            float *p = nullptr;
            CK(cudaMalloc(&p, sizeof(float) * size_t(cv->width) * size_t(cv->band_rows)));   // per canvas
            cv->masks[d.mask_dst] = p;
        }
    }
    sf.mask_table.assign(size_t(max_slot) + 1, nullptr);
    for (auto &kv : cv->masks)
        if (kv.first <= max_slot) sf.mask_table[kv.first] = kv.second;
    for (const draw_rec &d : sf.draws)
        if (d.mask_src && !sf.mask_table[d.mask_src])
            return fail(CB200_ERR_BAD_ARG, "draw reads a clip-mask slot that was never written");

    // glyph atlases: extend the device mirrors to cover what this frame's instances refer to
    sf.atlas_table.assign(sf.atlas_views.size(), atlas_dev());
    for (size_t a = 0; a < sf.atlas_views.size(); ++a) {
        const cb200_glyph_atlas &v = sf.atlas_views[a];
        cb200_canvas::atlas_mirror &m = cv->atlases[v.id];
        if (v.n_outlines > m.n_outlines || v.n_segs > m.n_segs || v.n_points > m.n_points) {
            for (uint32_t o = m.n_outlines; o < v.n_outlines; ++o) {      // new outlines: check them once
                const cb200_glyph_outline &ol = v.outlines[o];
                if (uint64_t(ol.first_point) + ol.n_points > v.n_points || uint64_t(ol.first_seg) + ol.n_segs > v.n_segs ||
                    ol.out_points != ol.n_contours + 3 * ol.n_segs)
                    return fail(CB200_ERR_BAD_ARG, "glyph outline out of range");
                for (uint32_t k = 0; k < ol.n_segs; ++k) {
                    const cb200_glyph_seg &sg = v.segs[ol.first_seg + k];
                    const uint32_t top = std::max(std::max(uint32_t(sg.from_a), uint32_t(sg.from_b)),
                                                  std::max(std::max(uint32_t(sg.to_a), uint32_t(sg.to_b)), uint32_t(sg.ctrl)));
                    if (top >= ol.n_points || sg.out < 1 || uint64_t(sg.out) + 3 > ol.out_points)
                        return fail(CB200_ERR_BAD_ARG, "glyph piece out of range");
                }
            }
            // a buffer that has to grow is re-created (cudaFree waits for frames in flight), then refilled
            const bool regrow = v.n_outlines > m.outlines.cap || v.n_segs > m.segs.cap || v.n_points > m.points.cap;
            if (regrow) {
                CK(m.outlines.reserve(std::max<size_t>(64, 2 * size_t(v.n_outlines))));
                CK(m.segs.reserve(std::max<size_t>(1024, 2 * size_t(v.n_segs))));
                CK(m.points.reserve(std::max<size_t>(1024, 2 * size_t(v.n_points))));
                m.n_outlines = m.n_segs = m.n_points = 0;
            }
            CK(cudaMemcpyAsync(m.outlines.p + m.n_outlines, v.outlines + m.n_outlines,
                               sizeof(cb200_glyph_outline) * (v.n_outlines - m.n_outlines), cudaMemcpyHostToDevice, cv->stream));
            CK(cudaMemcpyAsync(m.segs.p + m.n_segs, v.segs + m.n_segs, sizeof(cb200_glyph_seg) * (v.n_segs - m.n_segs),
                               cudaMemcpyHostToDevice, cv->stream));
            CK(cudaMemcpyAsync(m.points.p + m.n_points, v.points + 2 * size_t(m.n_points),
                               sizeof(float2) * (v.n_points - m.n_points), cudaMemcpyHostToDevice, cv->stream));
            m.n_outlines = v.n_outlines; m.n_segs = v.n_segs; m.n_points = v.n_points;
        }
        atlas_dev d = { m.outlines.p, m.segs.p, m.points.p };
        sf.atlas_table[a] = d;
    }

    std::vector<std::pair<size_t, std::pair<const void *, size_t> > > plan;
    size_t at = 0;
    std::vector<frame_header> hdr(1);
    memset(&hdr[0], 0, sizeof(frame_header));
    hdr[0].n_draws = uint32_t(sf.draws.size());
    hdr[0].n_subpaths = uint32_t(sf.subpaths.size());
    hdr[0].n_units = uint32_t(sf.units.size());
    hdr[0].n_jobs = uint32_t(sf.jobs.size());
    hdr[0].width = cv->width; hdr[0].height = cv->height;
    hdr[0].band_y0 = cv->band_y0; hdr[0].band_rows = cv->band_rows;
    hdr[0].n_sources = uint32_t(sf.sources.size());
    hdr[0].sort_bits_x = uint32_t(sf.bits_x); hdr[0].sort_bits_y = uint32_t(sf.bits_y);
    hdr[0].sort_bits = uint32_t(sf.key_bits);
    cv->hdr_offset = place(plan, at, hdr);
    cv->hdr_pristine_offset = place(plan, at, hdr);
    size_t o_draws = place(plan, at, sf.draws), o_sub = place(plan, at, sf.subpaths);
    size_t o_units = place(plan, at, sf.units);
    size_t o_brushes = place(plan, at, sf.brushes), o_colors = place(plan, at, sf.colors);
    size_t o_stops = place(plan, at, sf.stops), o_dashes = place(plan, at, sf.dashes);
    size_t o_ditems = place(plan, at, sf.dash_items), o_dsrc = place(plan, at, sf.draw_src);
    size_t o_jobs = place(plan, at, sf.jobs), o_sjobs = place(plan, at, sf.shadow_jobs);
    size_t o_masks = place(plan, at, sf.mask_table), o_tex = place(plan, at, sf.texels_u8);
    size_t o_cjobs = place(plan, at, sf.canvas_jobs);
    size_t o_src = place(plan, at, sf.sources);
    size_t o_ginst = place(plan, at, sf.glyph_insts), o_atlas = place(plan, at, sf.atlas_table);
    // points go last: the glyph-instance region follows them in the device blob and is never uploaded
    size_t o_points = place(plan, at, sf.points);
    const size_t upload_bytes = at;
    at = align_up(o_points + (sf.points.size() / 2 + size_t(sf.n_glyph_points)) * sizeof(float2) + 16, 256);

    if (at > cv->pinned_cap) {
        if (cv->pinned) cudaFreeHost(cv->pinned);
        cv->pinned = nullptr;
        cv->pinned_cap = 0;
        size_t want = align_up(at + at / 2, 1 << 16);
        CK(cudaMallocHost(&cv->pinned, want));
        cv->pinned_cap = want;
    }
    CK(cv->blob.reserve(cv->pinned_cap));
    for (auto &pl : plan)
        if (pl.second.second) memcpy(cv->pinned + pl.first, pl.second.first, pl.second.second);
    CK(cudaMemcpyAsync(cv->blob.p, cv->pinned, upload_bytes, cudaMemcpyHostToDevice, cv->stream));
    // static stroke sources live in the growable device array K2 appends to
    if (!sf.sources.empty())
        CK(cudaMemcpyAsync(cv->sources.p, cv->blob.p + o_src, sf.sources.size() * sizeof(stroke_src),
                           cudaMemcpyDeviceToDevice, cv->stream));

    device_frame &f = cv->df;
    memset(&f, 0, sizeof f);
    uint8_t *b = cv->blob.p;
    f.hdr = reinterpret_cast<frame_header *>(b + cv->hdr_offset);
    f.draws = reinterpret_cast<draw_rec *>(b + o_draws);
    f.subpaths = reinterpret_cast<subpath_rec *>(b + o_sub);
    f.units = reinterpret_cast<unit_rec *>(b + o_units);
    f.in_points = reinterpret_cast<float2 *>(b + o_points);
    f.glyph_insts = reinterpret_cast<glyph_inst_rec *>(b + o_ginst);
    f.n_glyph_insts = uint32_t(sf.glyph_insts.size());
    f.atlas_table = reinterpret_cast<atlas_dev *>(b + o_atlas);
    f.brushes = reinterpret_cast<brush_rec *>(b + o_brushes);
    f.colors = reinterpret_cast<float4 *>(b + o_colors);
    f.stops = reinterpret_cast<float *>(b + o_stops);
    f.dashes = reinterpret_cast<float *>(b + o_dashes);
    f.dash_items = reinterpret_cast<dash_item *>(b + o_ditems);
    f.n_dash_items = uint32_t(sf.dash_items.size());
    f.draw_src = reinterpret_cast<uint2 *>(b + o_dsrc);
    f.sources = cv->sources.p;
    f.n_static_sources = uint32_t(sf.sources.size());
    f.jobs = reinterpret_cast<job_rec *>(b + o_jobs);
    f.comp = cv->comp.p; f.job_box = cv->job_box.p; f.job_te = cv->job_te.p; f.blur_units = cv->blur_units.p;
    f.row_jobs = cv->use_row_lists ? cv->row_jobs.p : nullptr; f.row_job_count = cv->row_job_count.p; f.row_stride = cv->row_stride;
    f.n_opaque_jobs = 0;
    for (const job_rec &j : sf.jobs) f.n_opaque_jobs += j.opaque;
    f.general_compositor = 0;
    bool needs_planes = false, needs_gradients = false, needs_patterns = false, needs_everything = false;
    for (const job_rec &j : sf.jobs) {
        const draw_rec &d = sf.draws[j.draw];
        const brush_rec &br = sf.brushes[d.brush];
        if (j.kind != JOB_MAIN || d.mask_src) needs_planes = true;
        if (j.kind == JOB_MAIN && br.type != CB200_BRUSH_COLOR) {
            if (br.type == CB200_BRUSH_PATTERN) needs_patterns = true;
            else if (br.n_colors > 16) needs_everything = true;                                // 16 = kStagedStops
            else needs_gradients = true;
        }
    }
    // 0 lean, 1 + masks / shadows / clips, 2 lean + small gradients, 3 everything, 4 lean + patterns + small gradients
    if (needs_everything || (needs_planes && (needs_gradients || needs_patterns))) f.general_compositor = 3;
    else if (needs_patterns) f.general_compositor = 4;
    else if (needs_gradients) f.general_compositor = 2;
    else if (needs_planes) f.general_compositor = 1;
    f.shadow_jobs = reinterpret_cast<uint32_t *>(b + o_sjobs);
    f.n_shadow_jobs = uint32_t(sf.shadow_jobs.size());
    f.max_shadow_pad = 0; f.max_shadow_radius = 0; f.min_shadow_radius = 1 << 30;
    for (uint32_t sj : sf.shadow_jobs) {
        f.min_shadow_radius = std::min(f.min_shadow_radius, int(sf.jobs[sj].radius));
        f.max_shadow_pad = std::max(f.max_shadow_pad, int(sf.jobs[sj].pad));
        f.max_shadow_radius = std::max(f.max_shadow_radius, int(sf.jobs[sj].radius));
    }
    f.texels = cv->texels.p;
    f.unit_count = cv->unit_count.p; f.unit_offset = cv->unit_offset.p;
    f.pts = cv->pts.p; f.cap_pts = cv->cap_pts; f.pt_loop = cv->pt_loop.p;
    f.loops = cv->loops.p; f.cap_loops = uint32_t(cv->loops.cap);
    f.dash_pts_count = cv->dash_pts_count.p; f.dash_sub_count = cv->dash_sub_count.p; f.dash_tail = cv->dash_tail.p;
    f.half_count = cv->half_count.p; f.half_offset = cv->half_offset.p;
    f.half_unit_off = cv->half_unit_off.p; f.half_dirty = cv->half_dirty.p; f.stroke_unit_pts = cv->stroke_unit_pts.p;
    f.half_last = cv->half_last.p; f.visit_prev = cv->visit_prev.p; f.visit_close = cv->visit_close.p;
    f.cap_stroke_units = uint32_t(cv->stroke_unit_pts.cap);
    f.cap_sources = cv->cap_sources;
    f.stroke_loop_base = uint32_t(sf.subpaths.size()) + cv->cap_dash_subpaths;
    f.pieces = cv->pieces.p; f.piece_job = cv->piece_job.p; f.piece_rows = cv->piece_rows.p;
    f.piece_rlo = cv->piece_rlo.p; f.piece_row_off = cv->piece_row_off.p;
    f.cap_items = cv->cap_items;
    f.row_runs = cv->row_runs.p; f.row_piece = cv->row_piece.p; f.cap_rows = cv->cap_rows;
    f.keys[0] = cv->keys0.p; f.keys[1] = cv->keys1.p; f.vals[0] = cv->vals0.p; f.vals[1] = cv->vals1.p;
    f.cap_runs = cv->cap_runs; f.cumulative = cv->cumulative.p; f.long_rows = cv->long_rows.p;
    f.loop_mark = cv->loop_mark.p; f.box_loops = cv->box_loops.p;
    f.leaks = cv->leaks.p; f.cap_leaks = cv->cap_leaks;
    f.tile_cover = cv->tile_cover.p; f.n_target_tiles = uint32_t(cv->tile_cover.cap);
    f.te_flags = cv->te_flags.p; f.te_job = cv->te_job.p; f.te_backdrop = cv->te_backdrop.p; f.te_first = cv->te_first.p; f.te_mask = cv->te_mask.p;
    f.cap_tiles = cv->cap_tiles;
    f.planes = cv->planes.p; f.planes_tmp = cv->planes_tmp.p; f.cap_planes = cv->cap_planes;
    f.partials = cv->partials.p; f.sort_hist = cv->sort_hist.p;
    // many small jobs (batches): every job's runs are sorted inside one CTA instead of by the global passes (sort.cu)
    f.job_run_begin = segmented_sort_passes(sf.key_bits, sf.bits_x, sf.bits_y, sf.jobs.size()) ? cv->job_run_begin.p : nullptr;

    canvas_target &t = cv->target;
    t.fb = cv->fb; t.width = cv->width; t.height = cv->height;
    t.band_y0 = cv->band_y0; t.band_rows = cv->band_rows;
    t.mask_planes = reinterpret_cast<float **>(b + o_masks);
    t.n_masks = uint32_t(sf.mask_table.size());
    t.n_canvases = cv->n_canvases;
    t.slot_rows = cv->n_canvases > 1 ? cv->slot_rows : cv->band_rows;
    t.canvas_jobs = reinterpret_cast<uint2 *>(b + o_cjobs);

    if (sf.n_texels) {
        launch_texel_convert(b + o_tex, cv->texels.p, sf.n_texels, cv->stream);
        ++cv->launches;
    }
    return CB200_OK;
}

int frame_launch_count(const cb200_canvas *cv)
{
    const staged_frame &sf = cv->staged;
    const device_frame &f = cv->df;
    return (sf.glyph_insts.empty() ? 0 : 1) + (sf.units.empty() ? 0 : 3) + (sf.dash_items.empty() ? 0 : 2) +
           ((sf.sources.empty() && sf.dash_items.empty()) ? 0 : 9) + 7 + 3 * sort_passes(sf.key_bits) + (f.job_run_begin ? 2 : 0) + 3 +
           (sf.shadow_jobs.empty() ? 0 : 1 + (f.min_shadow_radius <= 30 ? 3 : 0) + (f.max_shadow_radius > 30 ? 5 : 0)) + 1 + (f.row_jobs ? 1 : 0);
}

// The fixed launch sequence of one frame, issued into the canvas stream -- or, with `in_graph`, into a
// stream capture: events the host later waits on or times become event-record nodes, the per-frame
// compositor event ring is skipped (a graph's nodes are the same for every replay).
int enqueue_frame(cb200_canvas *cv, bool in_graph)
{
    staged_frame &sf = cv->staged;
    device_frame &f = cv->df;
    cudaStream_t s = cv->stream;
    auto mark = [&](cudaEvent_t e) { return in_graph ? cudaEventRecordWithFlags(e, s, cudaEventRecordExternal) : cudaEventRecord(e, s); };
    CK(cudaMemsetAsync(cv->partials.p, 0, sizeof(uint32_t) * 8 * kGrid, s));
    if (f.n_shadow_jobs) CK(cudaMemsetAsync(cv->loop_mark.p, 0, sizeof(uint32_t) * f.cap_loops, s));
    CK(cudaMemsetAsync(cv->tile_cover.p, 0, sizeof(uint32_t) * f.n_target_tiles, s));
    // Stage events sit between kernels and so cut the programmatic-dependent-launch chain there;
    // with stage timing off only the frame and the compositor are bracketed.
    const bool stages = cv->stage_timing && !in_graph;
    CK(mark(cv->ev[0]));
    launch_glyphs(f, s);
    launch_flatten(f, uint32_t(sf.units.size()), s);
    launch_dash(f, s);
    launch_stroke(f, s);
    if (stages) CK(cudaEventRecord(cv->ev[1], s));
    launch_raster(f, cv->target, s);
    if (stages) CK(cudaEventRecord(cv->ev[2], s));
    int sorted = 0;
    launch_sort(f, s, sf.key_bits, &sorted, segmented_sort_passes(sf.key_bits, sf.bits_x, sf.bits_y, sf.jobs.size()),
                sf.bits_x + sf.bits_y, uint32_t(sf.jobs.size()));
    if (stages) CK(cudaEventRecord(cv->ev[3], s));
    launch_rows(f, cv->target, sorted, s);
    if (stages) CK(cudaEventRecord(cv->ev[8], s));
    launch_shadow(f, cv->target, sorted, s, stages ? cv->ev[9] : nullptr);
    CK(mark(cv->ev[4]));
    const int ring = int(cv->frames_run % uint64_t(cb200_canvas::kCompRing));
    if (!in_graph) CK(cudaEventRecord(cv->comp_ev[2 * ring], s));
    cv->target.clear_first = cv->inflight_clear ? 1 : 0;
    launch_composite(f, cv->target, sorted, s);
    if (!in_graph) CK(cudaEventRecord(cv->comp_ev[2 * ring + 1], s));
    CK(mark(cv->ev[5]));
    CK(cudaMemcpyAsync(cv->pinned_hdr, f.hdr, sizeof(frame_header), cudaMemcpyDeviceToHost, s));
    CK(mark(cv->ev[6]));
    return CB200_OK;
}

int run_frame(cb200_canvas *cv)
{
    int rc = enqueue_frame(cv, false);
    if (rc != CB200_OK) return rc;
    cv->comp_ev_valid[cv->frames_run % uint64_t(cb200_canvas::kCompRing)] = true;
    ++cv->frames_run;
    cv->launches += uint64_t(frame_launch_count(cv));
    cv->pending = true;
    CK(cudaGetLastError());
    return CB200_OK;
}

void drop_replay_graphs(cb200_canvas *cv)
{
    for (int k = 0; k < 2; ++k)
        if (cv->replay_graph[k]) { cudaGraphExecDestroy(cv->replay_graph[k]); cv->replay_graph[k] = nullptr; }
}

// One replay of the resident, verified frame as a single graph launch: header restore, every kernel
// of the frame (programmatic dependent launches become programmatic graph edges), header readback.
// Returns false when the graph path is not available (the caller then issues the stream sequence).
bool replay_with_graph(cb200_canvas *cv, int *rc_out)
{
    *rc_out = CB200_OK;
    if (!cv->graph_replay || cv->graph_broken || cv->stage_timing || !cv->replay_verified) return false;
    const int key = cv->inflight_clear ? 1 : 0;
    cudaStream_t s = cv->stream;
    if (!cv->replay_graph[key]) {
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
            ok = cudaMemcpyAsync(cv->blob.p + cv->hdr_offset, cv->blob.p + cv->hdr_pristine_offset, sizeof(frame_header),
                                 cudaMemcpyDeviceToDevice, s) == cudaSuccess;
            ok = ok && enqueue_frame(cv, true) == CB200_OK;
            ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && ok && graph;
        }
        if (ok) ok = cudaGraphInstantiate(&cv->replay_graph[key], graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();                          // clear the sticky-free error of the failed attempt
            cv->replay_graph[key] = nullptr;
            cv->graph_broken = true;
            if (getenv("CB200_DEBUG")) fprintf(stderr, "[cb200] graph capture of the frame failed; staying on stream launches\n");
            return false;
        }
    }
    cudaError_t e = cudaGraphLaunch(cv->replay_graph[key], s);
    if (e != cudaSuccess) { *rc_out = fail(CB200_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e)); return true; }
    cv->comp_ev_valid[cv->frames_run % uint64_t(cb200_canvas::kCompRing)] = false;
    ++cv->frames_run;
    ++cv->graph_replays;
    cv->launches += uint64_t(frame_launch_count(cv));
    cv->pending = true;
    return true;
}

// Wait for the launched frame, and if a device queue overflowed re-run it with
// larger buffers (the framebuffer was not touched by the failed attempt).
int finish_pending(cb200_canvas *cv)
{
    if (!cv->pending) return CB200_OK;
    for (int attempt = 0; attempt < 8; ++attempt) {
        CK(cudaEventSynchronize(cv->ev[6]));
        CK(cudaGetLastError());
        frame_header seen = *cv->pinned_hdr;
        if (getenv("CB200_DEBUG"))
            fprintf(stderr, "[cb200] attempt %d overflow=%#x line_pts=%u stroke_units=%u stroke_pts=%u items=%u rows=%u runs=%u "
                    "tiles=%u long=%u planes=%llu composited=%llu | caps pts=%u items=%u rows=%u runs=%u tiles=%u\n",
                    attempt, seen.overflow, seen.n_line_points, seen.n_stroke_units, seen.n_stroke_points, seen.n_items,
                    seen.n_row_items, seen.n_runs, seen.n_tile_entries, seen.n_long_rows,
                    (unsigned long long)seen.plane_floats, seen.composited_slots[0], cv->cap_pts, cv->cap_items, cv->cap_rows,
                    cv->cap_runs, cv->cap_tiles);
        if (getenv("CB200_DEBUG") && cv->staged.subpaths.size() < 64) {
            size_t ns = cv->staged.subpaths.size(), nu = cv->staged.units.size(), nsrc = cv->staged.sources.size();
            std::vector<loop_span> lp(ns + 1);
            std::vector<uint32_t> uo(nu + 1), uc(nu + 1);
            std::vector<stroke_src> ss(nsrc + 1);
            cudaMemcpy(lp.data(), cv->loops.p, ns * sizeof(loop_span), cudaMemcpyDeviceToHost);
            cudaMemcpy(uo.data(), cv->unit_offset.p, (nu + 1) * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(uc.data(), cv->unit_count.p, nu * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(ss.data(), cv->sources.p, nsrc * sizeof(stroke_src), cudaMemcpyDeviceToHost);
            fprintf(stderr, "  loops:");
            for (size_t i = 0; i < ns; ++i) fprintf(stderr, " (%u,%u)", lp[i].first, lp[i].count);
            fprintf(stderr, "\n  unit_offset:");
            for (size_t i = 0; i <= nu; ++i) fprintf(stderr, " %u", uo[i]);
            fprintf(stderr, "\n  unit_count:");
            for (size_t i = 0; i < nu; ++i) fprintf(stderr, " %u", uc[i]);
            fprintf(stderr, "\n  sources:");
            for (size_t i = 0; i < nsrc; ++i) fprintf(stderr, " (%u,%#x)", ss[i].loop, ss[i].draw_closed);
            fprintf(stderr, "\n");
        }
        if (!seen.overflow) {
            float ms = 0.0f;
            cb200_stats &st = cv->stats;
            cudaEventElapsedTime(&ms, cv->ev[0], cv->ev[5]); st.last_frame_ms = ms;
            cudaEventElapsedTime(&ms, cv->ev[4], cv->ev[5]); st.composite_ms = ms;
            st.geometry_ms = st.raster_ms = st.sort_ms = st.coverage_ms = st.shadow_raster_ms = st.blur_ms = 0.0f;
            if (cv->stage_timing) {
                cudaEventElapsedTime(&ms, cv->ev[0], cv->ev[1]); st.geometry_ms = ms;
                cudaEventElapsedTime(&ms, cv->ev[1], cv->ev[2]); st.raster_ms = ms;
                cudaEventElapsedTime(&ms, cv->ev[2], cv->ev[3]); st.sort_ms = ms;
                cudaEventElapsedTime(&ms, cv->ev[3], cv->ev[8]); st.coverage_ms = ms;
                cudaEventElapsedTime(&ms, cv->ev[8], cv->ev[9]); st.shadow_raster_ms = ms;
                cudaEventElapsedTime(&ms, cv->ev[9], cv->ev[4]); st.blur_ms = ms;
            }
            st.draws = seen.n_draws;
            st.cubics = seen.n_units - seen.n_subpaths;
            st.line_points = seen.n_line_points + seen.n_dash_points + seen.n_stroke_points;
            st.edges = seen.n_items;
            st.raw_runs = seen.n_runs;
            st.tile_entries = seen.n_tile_entries;
            st.composited_pixels = seen.composited_pixels;
            for (int k = 0; k < 32; ++k) st.composited_pixels += seen.composited_slots[k];
            st.shadow_pixels = seen.shadow_working_pixels;
            st.kernel_launches = cv->launches;
            st.graph_replays = cv->graph_replays;
            cv->pending = false;
            cv->replay_verified = cv->resident;
            return CB200_OK;
        }
        int rc = ensure_capacity(cv, cv->staged, &seen);
        if (rc != CB200_OK) return rc;
        // pointers/capacities changed: rebuild the device frame from the staged copy
        rc = upload_frame(cv);
        if (rc != CB200_OK) return rc;
        rc = run_frame(cv);
        if (rc != CB200_OK) return rc;
    }
    cv->pending = false;
    return fail(CB200_ERR_OVERFLOW, "device work queues still overflow after regrowth");
}

// A new frame: a pending clear rides along (the compositor starts from transparent black and
// writes every tile) instead of costing a separate pass over the framebuffer.
int start_frame(cb200_canvas *cv)
{
    cv->inflight_clear = cv->clear_pending;
    cv->clear_pending = false;
    return run_frame(cv);
}

// Before anything outside a frame reads or writes pixels: finish the frame in flight and apply a
// clear that no frame has absorbed.
int settle(cb200_canvas *cv)
{
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    if (cv->clear_pending) {
        CK(cudaMemsetAsync(cv->fb, 0, sizeof(float4) * size_t(cv->width) * fb_rows(cv), cv->stream));
        cv->clear_pending = false;
    }
    return CB200_OK;
}

// Clip-mask planes were freed: a resident frame's device mask table and replay graphs hold their addresses.
// cb200_frame_replay then reports "no frame uploaded" instead of touching freed memory.
void invalidate_resident(cb200_canvas *cv)
{
    bool uses_masks = false;
    for (const draw_rec &d : cv->staged.draws) uses_masks = uses_masks || d.mask_src || d.mask_dst;
    if (!cv->resident || !uses_masks) return;
    cv->resident = false;
    cv->replay_verified = false;
    drop_replay_graphs(cv);
}

}  // namespace

// ------------------------------------------------------------------- C ABI ----

extern "C" {

int cb200_set_stage_timing(cb200_canvas *cv, int on)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    cv->stage_timing = on != 0;
    return CB200_OK;
}

int cb200_abi_version(void) { return CB200_ABI_VERSION; }

int cb200_struct_size(int which)
{
    switch (which) {
    case 0: return int(sizeof(cb200_draw));
    case 1: return int(sizeof(cb200_subpath));
    case 2: return int(sizeof(cb200_brush));
    case 3: return int(sizeof(cb200_image));
    case 4: return int(sizeof(cb200_frame));
    case 5: return int(sizeof(cb200_glyph_seg));
    case 6: return int(sizeof(cb200_glyph_outline));
    case 7: return int(sizeof(cb200_glyph_atlas));
    case 8: return int(sizeof(cb200_glyph_inst));
    default: return -1;
    }
}

int cb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *cb200_last_error(void) { return g_error.c_str(); }

static thread_local int g_batch_n = 1;       // consumed by cb200_canvas_create_band (set by cb200_batch_create)

int cb200_canvas_create_band(int width, int height, int band_y0, int band_rows, int device, cb200_canvas **out)
{
    if (!out) return fail(CB200_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (width < 1 || height < 1 || width > 32768 || height > 32768)
        return fail(CB200_ERR_BAD_ARG, "canvas size must be 1..32768");
    if (band_y0 < 0 || band_rows < 1 || band_y0 + band_rows > height)
        return fail(CB200_ERR_BAD_ARG, "band outside the canvas");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CB200_ERR_NO_DEVICE, "no CUDA device available (this back end has no CPU path)");
    }
    if (device < 0 || device >= n) return fail(CB200_ERR_BAD_ARG, "no such CUDA device");
    CK(cudaSetDevice(device));
    cb200_canvas *cv = new cb200_canvas;
    cv->device = device; cv->width = width; cv->height = height;
    cv->band_y0 = band_y0; cv->band_rows = band_rows;
    cv->n_canvases = g_batch_n;
    cv->slot_rows = (height + kTile - 1) / kTile * kTile;
    memset(&cv->stats, 0, sizeof cv->stats);
    cudaError_t err = cudaStreamCreateWithFlags(&cv->stream, cudaStreamNonBlocking);
    size_t px = size_t(width) * fb_rows(cv);
    if (err == cudaSuccess) err = cudaMalloc(&cv->fb, px * sizeof(float4));
    if (err == cudaSuccess) err = cudaMemsetAsync(cv->fb, 0, px * sizeof(float4), cv->stream);
    if (err == cudaSuccess) err = cudaMallocHost(&cv->pinned_hdr, sizeof(frame_header));
    for (int i = 0; i < 10 && err == cudaSuccess; ++i) err = cudaEventCreate(&cv->ev[i]);
    for (int i = 0; i < 2 * cb200_canvas::kCompRing && err == cudaSuccess; ++i) err = cudaEventCreate(&cv->comp_ev[i]);
    for (int i = 0; i < 2 && err == cudaSuccess; ++i) err = cudaEventCreate(&cv->timer_ev[i]);
    if (err != cudaSuccess) {
        std::string why = cudaGetErrorString(err);
        cb200_canvas_destroy(cv);
        return fail(err == cudaErrorMemoryAllocation ? CB200_ERR_OOM : CB200_ERR_CUDA, "canvas create: " + why);
    }
    *out = cv;
    return CB200_OK;
}

int cb200_batch_create(int n_canvases, int width, int height, int device, cb200_canvas **out)
{
    if (n_canvases < 1 || n_canvases > (1 << 22)) return fail(CB200_ERR_BAD_ARG, "batch size must be 1..4194304");
    g_batch_n = n_canvases;
    int rc = cb200_canvas_create_band(width, height, 0, height, device, out);
    g_batch_n = 1;
    return rc;
}

int cb200_batch_submit(cb200_canvas *cv, const cb200_frame *const *frames, const uint32_t *canvas_index,
                       uint32_t n_frames)
{
    if (!cv || (n_frames && (!frames || !canvas_index))) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    if (n_frames == 0) return CB200_OK;
    rc = stage_frames(cv, frames, canvas_index, n_frames, cv->staged);
    if (rc != CB200_OK) return rc;
    if (cv->staged.draws.empty()) return CB200_OK;
    cv->resident = false;
    rc = ensure_capacity(cv, cv->staged, nullptr);
    if (rc != CB200_OK) return rc;
    rc = upload_frame(cv);
    if (rc != CB200_OK) return rc;
    return start_frame(cv);
}

int cb200_batch_read_rgba8(cb200_canvas *cv, uint32_t canvas, uint8_t *dst, int width, int height, int stride,
                           int x, int y)
{
    if (!cv || !dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (canvas >= uint32_t(cv->n_canvases)) return fail(CB200_ERR_BAD_ARG, "canvas index out of range");
    if (width <= 0 || height <= 0) return CB200_OK;
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    size_t bytes = 4 * size_t(width) * size_t(height);
    CK(cv->rgba8.reserve(std::max<size_t>(bytes, 16)));
    const float4 *slot = cv->fb + size_t(canvas) * size_t(cv->n_canvases > 1 ? cv->slot_rows : 0) * size_t(cv->width);
    launch_readback(slot, cv->width, 0, cv->height, cv->rgba8.p, width, height, x, y, cv->stream);
    ++cv->launches;
    std::vector<uint8_t> tmp(bytes);
    CK(cudaMemcpyAsync(tmp.data(), cv->rgba8.p, bytes, cudaMemcpyDeviceToHost, cv->stream));
    CK(cudaStreamSynchronize(cv->stream));
    for (int row = 0; row < height; ++row)
        memcpy(dst + ptrdiff_t(row) * stride, tmp.data() + size_t(row) * size_t(width) * 4, size_t(width) * 4);
    return CB200_OK;
}

int cb200_batch_write_rgba8(cb200_canvas *cv, uint32_t canvas, const uint8_t *src, int width, int height, int stride,
                            int x, int y)
{
    if (!cv || !src) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (canvas >= uint32_t(cv->n_canvases)) return fail(CB200_ERR_BAD_ARG, "canvas index out of range");
    if (width <= 0 || height <= 0) return CB200_OK;
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    const size_t bytes = 4 * size_t(width) * size_t(height);
    std::vector<uint8_t> packed(bytes);
    for (int row = 0; row < height; ++row)
        memcpy(packed.data() + size_t(row) * size_t(width) * 4, src + ptrdiff_t(row) * stride, size_t(width) * 4);
    CK(cv->rgba8.reserve(std::max<size_t>(bytes, 16)));
    CK(cudaMemcpyAsync(cv->rgba8.p, packed.data(), bytes, cudaMemcpyHostToDevice, cv->stream));
    float4 *slot = cv->fb + size_t(canvas) * size_t(cv->n_canvases > 1 ? cv->slot_rows : 0) * size_t(cv->width);
    launch_upload(slot, cv->width, 0, cv->height, cv->rgba8.p, width, height, x, y, cv->stream);
    ++cv->launches;
    CK(cudaStreamSynchronize(cv->stream));       // `packed` is pageable: the copy above has completed, the kernel too
    return CB200_OK;
}

int cb200_batch_masks_keep(cb200_canvas *cv, const uint32_t *canvas, const uint32_t *local_slot, uint32_t n)
{
    if (!cv || (n && (!canvas || !local_slot))) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaStreamSynchronize(cv->stream));
    // batch-wide slots that stay: the ones the listed (canvas, local slot) pairs map to
    std::vector<uint32_t> keep;
    for (auto ci = cv->batch_masks.begin(); ci != cv->batch_masks.end();) {
        for (auto si = ci->second.begin(); si != ci->second.end();) {
            bool alive = false;
            for (uint32_t i = 0; i < n; ++i) alive = alive || (canvas[i] == ci->first && local_slot[i] == si->first);
            if (alive) { keep.push_back(si->second); ++si; }
            else si = ci->second.erase(si);
        }
        if (ci->second.empty()) ci = cv->batch_masks.erase(ci); else ++ci;
    }
    for (auto it = cv->masks.begin(); it != cv->masks.end();) {
        if (std::find(keep.begin(), keep.end(), it->first) != keep.end()) ++it;
        else { cudaFree(it->second); it = cv->masks.erase(it); }
    }
    cv->resident = false;                                   // a resident frame's mask table may name freed planes
    cv->replay_verified = false;
    drop_replay_graphs(cv);
    return CB200_OK;
}

int cb200_batch_read_f32(cb200_canvas *cv, uint32_t canvas, float *dst)
{
    if (!cv || !dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (canvas >= uint32_t(cv->n_canvases)) return fail(CB200_ERR_BAD_ARG, "canvas index out of range");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    const float4 *slot = cv->fb + size_t(canvas) * size_t(cv->n_canvases > 1 ? cv->slot_rows : 0) * size_t(cv->width);
    CK(cudaMemcpyAsync(dst, slot, sizeof(float4) * size_t(cv->width) * size_t(cv->height), cudaMemcpyDeviceToHost, cv->stream));
    CK(cudaStreamSynchronize(cv->stream));
    return CB200_OK;
}

int cb200_canvas_create(int width, int height, int device, cb200_canvas **out)
{
    return cb200_canvas_create_band(width, height, 0, height, device, out);
}

void cb200_canvas_destroy(cb200_canvas *cv)
{
    if (!cv) return;
    cudaSetDevice(cv->device);
    if (cv->stream) cudaStreamSynchronize(cv->stream);
    if (cv->fb) cudaFree(cv->fb);
    for (auto &kv : cv->masks) cudaFree(kv.second);
    cv->blob.release();
    if (cv->pinned) cudaFreeHost(cv->pinned);
    if (cv->pinned_hdr) cudaFreeHost(cv->pinned_hdr);
    if (cv->pinned_rgba8) cudaFreeHost(cv->pinned_rgba8);
    cv->unit_count.release(); cv->unit_offset.release(); cv->pt_loop.release();
    cv->dash_pts_count.release(); cv->dash_sub_count.release(); cv->dash_tail.release();
    cv->half_count.release(); cv->half_offset.release(); cv->half_unit_off.release(); cv->half_dirty.release(); cv->stroke_unit_pts.release(); cv->half_last.release(); cv->visit_prev.release(); cv->visit_close.release(); cv->piece_job.release();
    cv->piece_rows.release(); cv->piece_rlo.release(); cv->piece_row_off.release();
    cv->row_runs.release(); cv->row_piece.release(); cv->te_flags.release(); cv->te_job.release(); cv->te_first.release(); cv->te_mask.release(); cv->partials.release();
    cv->sort_hist.release(); cv->pts.release(); cv->loops.release(); cv->sources.release();
    cv->pieces.release(); cv->texels.release(); cv->comp.release(); cv->job_box.release(); cv->job_te.release(); cv->job_run_begin.release(); cv->blur_units.release(); cv->row_jobs.release(); cv->row_job_count.release(); cv->keys0.release(); cv->keys1.release();
    cv->vals0.release(); cv->vals1.release(); cv->cumulative.release(); cv->long_rows.release(); cv->te_backdrop.release();
    cv->planes.release(); cv->planes_tmp.release(); cv->rgba8.release(); cv->loop_mark.release(); cv->box_loops.release(); cv->leaks.release(); cv->tile_cover.release();
    drop_replay_graphs(cv);
    cv->png_tables.release(); cv->png_row_crc.release(); cv->png_out.release(); cv->png_acc.release();
    cv->hit_edges.release(); cv->hit_queries.release(); cv->hit_acc.release(); cv->hit_inside.release();
    for (auto &kv : cv->atlases) { kv.second.outlines.release(); kv.second.segs.release(); kv.second.points.release(); }
    for (int i = 0; i < 10; ++i)
        if (cv->ev[i]) cudaEventDestroy(cv->ev[i]);
    for (cudaEvent_t e : cv->comp_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : cv->timer_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : cv->chunk_events) cudaEventDestroy(e);
    if (cv->stream) cudaStreamDestroy(cv->stream);
    delete cv;
}

int cb200_submit(cb200_canvas *cv, const cb200_frame *frame)
{
    if (!cv || !frame) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);                 // the previous frame owns the staging until verified
    if (rc != CB200_OK) return rc;
    if (frame->n_draws == 0) return CB200_OK;
    rc = stage_frame(cv, frame, cv->staged);
    if (rc != CB200_OK) return rc;
    cv->resident = false;
    rc = ensure_capacity(cv, cv->staged, nullptr);
    if (rc != CB200_OK) return rc;
    rc = upload_frame(cv);
    if (rc != CB200_OK) return rc;
    return start_frame(cv);
}

int cb200_frame_upload(cb200_canvas *cv, const cb200_frame *frame)
{
    if (!cv || !frame) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    rc = stage_frame(cv, frame, cv->staged);
    if (rc != CB200_OK) return rc;
    rc = ensure_capacity(cv, cv->staged, nullptr);
    if (rc != CB200_OK) return rc;
    rc = upload_frame(cv);
    if (rc != CB200_OK) return rc;
    cv->resident = true;
    cv->replay_verified = false;
    return CB200_OK;
}

int cb200_frame_keep(cb200_canvas *cv)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);                 // completed (and re-run with larger buffers if it had to be)
    if (rc != CB200_OK) return rc;
    if (!cv->staged.valid || cv->staged.draws.empty()) return fail(CB200_ERR_BAD_ARG, "no frame submitted");
    cv->resident = true;
    cv->replay_verified = false;
    return CB200_OK;
}

int cb200_frame_replay(cb200_canvas *cv, int clear)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (!cv->resident || !cv->staged.valid) return fail(CB200_ERR_BAD_ARG, "no frame uploaded");
    CK(cudaSetDevice(cv->device));
    // A frame that has already completed once with the present capacities cannot overflow, so
    // its replays are queued back to back without waiting for the previous one's header.
    if (!(cv->pending && cv->replay_verified)) {
        int rc = finish_pending(cv);
        if (rc != CB200_OK) return rc;
    }
    if (clear) cv->clear_pending = true;
    {
        const bool keep_clear = cv->clear_pending;
        cv->inflight_clear = cv->clear_pending;          // what start_frame() would set; the graph is keyed on it
        cv->clear_pending = false;
        int rc = CB200_OK;
        if (replay_with_graph(cv, &rc)) return rc;
        cv->clear_pending = keep_clear;
    }
    // the header is consumed by a run: restore it from its pristine twin (device to device)
    CK(cudaMemcpyAsync(cv->blob.p + cv->hdr_offset, cv->blob.p + cv->hdr_pristine_offset, sizeof(frame_header),
                       cudaMemcpyDeviceToDevice, cv->stream));
    return start_frame(cv);
}

int cb200_set_graph_replay(cb200_canvas *cv, int on)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    cv->graph_replay = on != 0;
    return CB200_OK;
}

int cb200_timer_begin(cb200_canvas *cv)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    CK(cudaEventRecord(cv->timer_ev[0], cv->stream));
    cv->timer_frame0 = cv->frames_run;
    return CB200_OK;
}

int cb200_timer_stop(cb200_canvas *cv)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    CK(cudaEventRecord(cv->timer_ev[1], cv->stream));
    cv->timer_stopped = true;
    return CB200_OK;
}

int cb200_timer_end(cb200_canvas *cv, float *elapsed_ms, float *composite_ms, uint32_t *composite_frames)
{
    if (!cv || !elapsed_ms) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    if (!cv->timer_stopped) CK(cudaEventRecord(cv->timer_ev[1], cv->stream));
    cv->timer_stopped = false;
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaEventSynchronize(cv->timer_ev[1]));
    CK(cudaEventElapsedTime(elapsed_ms, cv->timer_ev[0], cv->timer_ev[1]));
    uint64_t first = cv->timer_frame0;
    if (cv->frames_run - first > uint64_t(cb200_canvas::kCompRing)) first = cv->frames_run - cb200_canvas::kCompRing;
    float sum = 0.0f;
    uint32_t counted = 0;
    for (uint64_t k = first; k < cv->frames_run; ++k) {
        const int ring = int(k % uint64_t(cb200_canvas::kCompRing));
        if (!cv->comp_ev_valid[ring]) continue;            // a graph replay: no per-frame events
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, cv->comp_ev[2 * ring], cv->comp_ev[2 * ring + 1]));
        sum += ms;
        ++counted;
    }
    if (composite_ms) *composite_ms = sum;
    if (composite_frames) *composite_frames = counted;
    return CB200_OK;
}

void *cb200_stream(cb200_canvas *cv) { return cv ? static_cast<void *>(cv->stream) : nullptr; }

int cb200_timer_between(cb200_canvas *from, cb200_canvas *to, float *elapsed_ms)
{
    if (!from || !to || !elapsed_ms) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (from->device != to->device) return fail(CB200_ERR_BAD_ARG, "cb200_timer_between: canvases on different devices");
    CK(cudaSetDevice(from->device));
    CK(cudaEventSynchronize(to->timer_ev[1]));
    CK(cudaEventElapsedTime(elapsed_ms, from->timer_ev[0], to->timer_ev[1]));
    return CB200_OK;
}

int cb200_sync(cb200_canvas *cv)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaStreamSynchronize(cv->stream));
    return CB200_OK;
}

static int readback_to_device(cb200_canvas *cv, int width, int height, int x, int y)
{
    size_t bytes = 4 * size_t(width) * size_t(height);
    CK(cv->rgba8.reserve(std::max<size_t>(bytes, 16)));
    CK(cudaEventRecord(cv->ev[7], cv->stream));
    launch_readback(cv->fb, cv->width, cv->band_y0, cv->band_rows, cv->rgba8.p, width, height, x, y, cv->stream,
                    cv->read_bgra ? 1 : 0);
    ++cv->launches;
    return CB200_OK;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

void *cb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void cb200_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int cb200_read_bgra8(cb200_canvas *cv, uint8_t *dst, int width, int height, int stride, int x, int y)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    cv->read_bgra = true;
    int rc = cb200_read_rgba8(cv, dst, width, height, stride, x, y);
    cv->read_bgra = false;
    return rc;
}

int cb200_framebuffer_device(cb200_canvas *cv, void **device_ptr, int *rows, int *width)
{
    if (!cv || !device_ptr) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaStreamSynchronize(cv->stream));
    *device_ptr = cv->fb;
    if (rows) *rows = int(fb_rows(cv));
    if (width) *width = cv->width;
    return CB200_OK;
}

int cb200_read_rgba8(cb200_canvas *cv, uint8_t *dst, int width, int height, int stride, int x, int y)
{
    if (!cv || !dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (width <= 0 || height <= 0) return CB200_OK;
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    rc = readback_to_device(cv, width, height, x, y);
    if (rc != CB200_OK) return rc;
    const size_t row_bytes = 4 * size_t(width), bytes = row_bytes * size_t(height);
    if (stride >= int(row_bytes) && is_pinned_host(dst)) {
        // caller's buffer is page-locked: one strided DMA, no staging
        CK(cudaMemcpy2DAsync(dst, size_t(stride), cv->rgba8.p, row_bytes, row_bytes, size_t(height),
                             cudaMemcpyDeviceToHost, cv->stream));
        CK(cudaEventRecord(cv->ev[6], cv->stream));
        CK(cudaStreamSynchronize(cv->stream));
    } else {
        // pageable destination: DMA into pinned staging in chunks and copy each chunk out while the
        // next one is still in flight
        if (bytes > cv->pinned_rgba8_cap) {
            if (cv->pinned_rgba8) cudaFreeHost(cv->pinned_rgba8);
            cv->pinned_rgba8 = nullptr;
            cv->pinned_rgba8_cap = 0;
            CK(cudaMallocHost(&cv->pinned_rgba8, bytes));
            cv->pinned_rgba8_cap = bytes;
        }
        const int rows_per_chunk = std::max(1, int((size_t(4) << 20) / row_bytes));
        const int n_chunks = (height + rows_per_chunk - 1) / rows_per_chunk;
        if (int(cv->chunk_events.size()) < n_chunks) {
            size_t old = cv->chunk_events.size();
            cv->chunk_events.resize(size_t(n_chunks));
            for (size_t i = old; i < cv->chunk_events.size(); ++i)
                CK(cudaEventCreateWithFlags(&cv->chunk_events[i], cudaEventDisableTiming));
        }
        for (int c = 0; c < n_chunks; ++c) {
            int r0 = c * rows_per_chunk, rows = std::min(rows_per_chunk, height - r0);
            CK(cudaMemcpyAsync(cv->pinned_rgba8 + size_t(r0) * row_bytes, cv->rgba8.p + size_t(r0) * row_bytes,
                               size_t(rows) * row_bytes, cudaMemcpyDeviceToHost, cv->stream));
            CK(cudaEventRecord(cv->chunk_events[size_t(c)], cv->stream));
        }
        CK(cudaEventRecord(cv->ev[6], cv->stream));
        for (int c = 0; c < n_chunks; ++c) {
            int r0 = c * rows_per_chunk, rows = std::min(rows_per_chunk, height - r0);
            CK(cudaEventSynchronize(cv->chunk_events[size_t(c)]));
            if (stride == int(row_bytes))
                memcpy(dst + size_t(r0) * row_bytes, cv->pinned_rgba8 + size_t(r0) * row_bytes, size_t(rows) * row_bytes);
            else
                for (int row = r0; row < r0 + rows; ++row)
                    memcpy(dst + ptrdiff_t(row) * stride, cv->pinned_rgba8 + size_t(row) * row_bytes, row_bytes);
        }
    }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, cv->ev[7], cv->ev[6]);
    cv->stats.readback_ms = ms;
    return CB200_OK;
}

int cb200_encode_png(cb200_canvas *cv, uint8_t *dst, size_t capacity, size_t *bytes)
{
    if (!cv || (!dst && !bytes)) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (cv->n_canvases != 1 || cv->band_y0 != 0 || cv->band_rows != cv->height)
        return fail(CB200_ERR_BAD_ARG, "cb200_encode_png: whole canvases only");
    if (4 * size_t(cv->width) + 1 > 65535) return fail(CB200_ERR_BAD_ARG, "cb200_encode_png: a row must fit one stored deflate block (width <= 16383)");
    const size_t size = 76 + size_t(cv->height) * (6 + 4 * size_t(cv->width));
    if (bytes) *bytes = size;
    if (!dst) return CB200_OK;
    if (capacity < size) return fail(CB200_ERR_BAD_ARG, "cb200_encode_png: destination too small");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    cudaStream_t s = cv->stream;
    if (!cv->png_tables_ready) {
        CK(cv->png_tables.reserve(png_table_words(cv->width, cv->height)));
        CK(cv->png_acc.reserve(1));
        CK(cv->png_row_crc.reserve(size_t(cv->height)));
        CK(cv->png_out.reserve(size));
        launch_png_tables(cv->png_tables.p, cv->width, cv->height, s);
        ++cv->launches;
        cv->png_tables_ready = true;
    }
    CK(cudaEventRecord(cv->ev[7], s));
    launch_png_encode(cv->fb, cv->width, cv->height, cv->png_out.p, cv->png_tables.p, cv->png_acc.p, cv->png_row_crc.p, s);
    cv->launches += 2;
    CK(cudaEventRecord(cv->ev[6], s));
    if (is_pinned_host(dst)) CK(cudaMemcpyAsync(dst, cv->png_out.p, size, cudaMemcpyDeviceToHost, s));
    else {
        if (size > cv->pinned_rgba8_cap) {
            if (cv->pinned_rgba8) cudaFreeHost(cv->pinned_rgba8);
            cv->pinned_rgba8 = nullptr;
            cv->pinned_rgba8_cap = 0;
            CK(cudaMallocHost(&cv->pinned_rgba8, size));
            cv->pinned_rgba8_cap = size;
        }
        CK(cudaMemcpyAsync(cv->pinned_rgba8, cv->png_out.p, size, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (!is_pinned_host(dst)) memcpy(dst, cv->pinned_rgba8, size);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, cv->ev[7], cv->ev[6]);
    cv->stats.png_ms = ms;
    return CB200_OK;
}

int cb200_read_rgba8_device(cb200_canvas *cv, void **device_ptr)
{
    if (!cv || !device_ptr) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    rc = readback_to_device(cv, cv->width, cv->band_rows, 0, cv->band_y0);
    if (rc != CB200_OK) return rc;
    CK(cudaEventRecord(cv->ev[6], cv->stream));
    CK(cudaEventSynchronize(cv->ev[6]));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, cv->ev[7], cv->ev[6]);
    cv->stats.readback_ms = ms;                           // the sRGB + dither kernel alone on this path
    *device_ptr = cv->rgba8.p;
    return CB200_OK;
}

int cb200_read_rgba8_into(cb200_canvas *cv, void *device_dst, int width, int height, int x, int y)
{
    if (!cv || !device_dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (width <= 0 || height <= 0) return CB200_OK;
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    launch_readback(cv->fb, cv->width, cv->band_y0, cv->band_rows, static_cast<uint8_t *>(device_dst), width, height,
                    x, y, cv->stream);
    ++cv->launches;
    CK(cudaGetLastError());
    return CB200_OK;
}

int cb200_write_rgba8(cb200_canvas *cv, const uint8_t *src, int width, int height, int stride, int x, int y)
{
    if (!cv || !src) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (width <= 0 || height <= 0) return CB200_OK;
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    size_t bytes = 4 * size_t(width) * size_t(height);
    if (bytes > cv->pinned_rgba8_cap) {
        if (cv->pinned_rgba8) cudaFreeHost(cv->pinned_rgba8);
        cv->pinned_rgba8 = nullptr;
        cv->pinned_rgba8_cap = 0;
        CK(cudaMallocHost(&cv->pinned_rgba8, bytes));
        cv->pinned_rgba8_cap = bytes;
    }
    for (int row = 0; row < height; ++row)
        memcpy(cv->pinned_rgba8 + size_t(row) * size_t(width) * 4, src + ptrdiff_t(row) * stride, size_t(width) * 4);
    CK(cv->rgba8.reserve(std::max<size_t>(bytes, 16)));
    CK(cudaMemcpyAsync(cv->rgba8.p, cv->pinned_rgba8, bytes, cudaMemcpyHostToDevice, cv->stream));
    launch_upload(cv->fb, cv->width, cv->band_y0, cv->band_rows, cv->rgba8.p, width, height, x, y, cv->stream);
    ++cv->launches;
    CK(cudaStreamSynchronize(cv->stream));       // pinned staging is reused by the next call
    return CB200_OK;
}

int cb200_read_f32(cb200_canvas *cv, float *dst)
{
    if (!cv || !dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = settle(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaMemcpyAsync(dst, cv->fb, sizeof(float4) * size_t(cv->width) * size_t(cv->band_rows),
                       cudaMemcpyDeviceToHost, cv->stream));
    CK(cudaStreamSynchronize(cv->stream));
    return CB200_OK;
}

int cb200_read_mask(cb200_canvas *cv, uint32_t slot, float *dst)
{
    if (!cv || !dst) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    size_t n = size_t(cv->width) * size_t(cv->band_rows);
    if (slot == 0) { for (size_t i = 0; i < n; ++i) dst[i] = 1.0f; return CB200_OK; }
    if (!cv->masks.count(slot)) return fail(CB200_ERR_BAD_ARG, "no such mask slot");
    CK(cudaMemcpyAsync(dst, cv->masks[slot], sizeof(float) * n, cudaMemcpyDeviceToHost, cv->stream));
    CK(cudaStreamSynchronize(cv->stream));
    return CB200_OK;
}

int cb200_hit_test(cb200_canvas *cv, const float *edges, uint32_t n_edges, const float *queries,
                   uint32_t n_queries, uint8_t *inside, float *kernel_ms)
{
    if (!cv || (n_edges && !edges) || (n_queries && (!queries || !inside))) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    if (kernel_ms) *kernel_ms = 0.0f;
    if (!n_queries) return CB200_OK;
    CK(cv->hit_edges.reserve(std::max<size_t>(n_edges, 1)));
    CK(cv->hit_queries.reserve(n_queries));
    CK(cv->hit_acc.reserve(n_queries));
    CK(cv->hit_inside.reserve(n_queries));
    cudaStream_t s = cv->stream;
    if (n_edges) CK(cudaMemcpyAsync(cv->hit_edges.p, edges, sizeof(float4) * size_t(n_edges), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(cv->hit_queries.p, queries, sizeof(float2) * size_t(n_queries), cudaMemcpyHostToDevice, s));
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (kernel_ms) { CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventRecord(t0, s)); }
    launch_hit_test(cv->hit_edges.p, n_edges, cv->hit_queries.p, n_queries, cv->hit_acc.p, cv->hit_inside.p, s);
    cv->launches += 2;
    if (kernel_ms) CK(cudaEventRecord(t1, s));
    CK(cudaMemcpyAsync(inside, cv->hit_inside.p, n_queries, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (kernel_ms) {
        CK(cudaEventElapsedTime(kernel_ms, t0, t1));
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
    }
    return CB200_OK;
}

int cb200_masks_keep(cb200_canvas *cv, const uint32_t *slots, uint32_t n)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    CK(cudaStreamSynchronize(cv->stream));
    bool freed = false;
    for (auto it = cv->masks.begin(); it != cv->masks.end();) {
        bool keep = false;
        for (uint32_t i = 0; i < n; ++i) keep = keep || slots[i] == it->first;
        if (keep) ++it;
        else { cudaFree(it->second); it = cv->masks.erase(it); freed = true; }
    }
    if (freed) invalidate_resident(cv);
    return CB200_OK;
}

int cb200_clear(cb200_canvas *cv)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    cv->clear_pending = true;                    // absorbed by the next frame, or applied by the next pixel access
    if (!cv->masks.empty()) {
        // The planes go (the caller resets its clip state with the canvas).  A resident frame holds raw plane
        // pointers in its uploaded mask table and in its captured graphs: it must be uploaded again.
        CK(cudaStreamSynchronize(cv->stream));
        for (auto &kv : cv->masks) cudaFree(kv.second);
        cv->masks.clear();
        cv->batch_masks.clear();
        invalidate_resident(cv);
    }
    return CB200_OK;
}

int cb200_get_stats(cb200_canvas *cv, cb200_stats *out)
{
    if (!cv || !out) return fail(CB200_ERR_BAD_ARG, "null argument");
    CK(cudaSetDevice(cv->device));
    int rc = finish_pending(cv);
    if (rc != CB200_OK) return rc;
    cv->stats.kernel_launches = cv->launches;
    *out = cv->stats;
    return CB200_OK;
}

// Host build of what the device enters into a shadow job's run bounding box for ONE closed loop of `n`
// points (device/edge_clip.cuh: clip_edge + the per-scanline add_runs of k_row_emit for the inside pieces,
// the boundary-segment walk of k_shadow_boxes for the rest).  box5 = (min_x, max_x, min_y, max_y, first run
// key y << 16 | x or -1); max < 0 when the loop enters nothing.  No device is touched: the CPU test feeds it
// the fuzz scenes' outlines and compares with the reference's polygon clip.
void cb200_debug_shadow_box(const float *xy, uint32_t n, float off_x, float off_y, int padded_w, int padded_h, int *box5)
{
    struct bounds_sink {
        int lo_x, hi_x;
        void put(float px, float delta) { if (delta != 0.0f) { lo_x = std::min(lo_x, int(px)); hi_x = std::max(hi_x, int(px)); } }
    };
    const float w = float(padded_w), h = float(padded_h);
    shadow_box_walk walk;
    walk.init(w, h);
    int lx = 0x7fffffff, ly = 0x7fffffff, hx = -1, hy = -1;
    long first_key = -1;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t q = k ? k - 1 : n - 1;
        clipped_edge ce;
        clip_edge(v2(off_x + xy[2 * q], off_y + xy[2 * q + 1]), v2(off_x + xy[2 * k], off_y + xy[2 * k + 1]), w, h, ce);
        for (int e = 0; e < ce.n_events; ++e) walk.consume(ce.ev[e].kind, ce.ev[e].v);
        for (int i = 0; i < ce.n_pieces; ++i) {
            const float4 pc = ce.piece[i];
            if (ce.projected[i] || !piece_has_runs(pc, false)) continue;
            const edge_walk ew = edge_setup(pc);
            for (int r = 0; r < ew.rows; ++r) {
                const row_walk rw = row_setup(ew, r);
                if (int(rw.py) < 0 || int(rw.py) >= padded_h) continue;        // k_edges keeps rows [0, padded_h)
                bounds_sink sink = { 0x7fffffff, -1 };
                walk_row_runs(ew, rw, sink);
                if (sink.hi_x >= 0) { lx = std::min(lx, sink.lo_x); hx = std::max(hx, sink.hi_x); ly = std::min(ly, int(rw.py)); hy = std::max(hy, int(rw.py)); }
                const long key = (long(rw.py) << 16) | long(rw.px);
                if (first_key < 0 || key < first_key) first_key = key;
            }
        }
    }
    walk.finish();
    if (walk.hx >= 0) { lx = std::min(lx, walk.lx); hx = std::max(hx, walk.hx); ly = std::min(ly, walk.ly); hy = std::max(hy, walk.hy); }
    box5[0] = lx; box5[1] = hx; box5[2] = ly; box5[3] = hy; box5[4] = int(first_key);
}

// Host build of the device's scan conversion of ONE closed loop (clip_edge + per-scanline add_runs, projected
// pieces included): every run as (x, y, delta), in edge / piece / scanline / pixel order.  Returns the run count
// and copies at most `capacity`.  The CPU test compares this multiset with the reference's own add_runs over the
// Sutherland-Hodgman-clipped loop, bit for bit (tests/test_scan_conversion.py).
int64_t cb200_debug_loop_runs(const float *xy, uint32_t n, float off_x, float off_y, int padded_w, int padded_h,
                              int32_t *run_xy, float *run_delta, int64_t capacity)
{
    struct collect_sink {
        int32_t *xy; float *delta; int64_t cap, count; int y;
        void put(float px, float d)
        {
            if (count < cap) { if (xy) { xy[2 * count] = int32_t(px); xy[2 * count + 1] = y; } if (delta) delta[count] = d; }
            ++count;
        }
    };
    collect_sink sink = { run_xy, run_delta, capacity, 0, 0 };
    const float w = float(padded_w), h = float(padded_h);
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t q = k ? k - 1 : n - 1;
        clipped_edge ce;
        clip_edge(v2(off_x + xy[2 * q], off_y + xy[2 * q + 1]), v2(off_x + xy[2 * k], off_y + xy[2 * k + 1]), w, h, ce);
        for (int i = 0; i < ce.n_pieces; ++i) {
            const float4 pc = ce.piece[i];
            if (!piece_has_runs(pc, ce.projected[i] != 0)) continue;
            const edge_walk ew = edge_setup(pc);
            for (int r = 0; r < ew.rows; ++r) {
                const row_walk rw = row_setup(ew, r);
                sink.y = int(rw.py);
                walk_row_runs(ew, rw, sink);
            }
        }
    }
    return sink.count;
}

// join_acosf / join_tanf (geom.cuh) over an array, on the host or on the device: the rounded join's two libm
// calls, which must carry the host libm's bits (tests/test_geometry_math.py compares with libm itself).
int cb200_debug_join_math(const float *x, uint32_t n, float *acos_out, float *tan_out, int on_device)
{
    if (!x || !acos_out || !tan_out) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (!on_device) {
        for (uint32_t i = 0; i < n; ++i) { acos_out[i] = join_acosf(x[i]); tan_out[i] = join_tanf(x[i]); }
        return CB200_OK;
    }
    float *dx = nullptr, *da = nullptr, *dt = nullptr;
    CK(cudaMalloc(&dx, sizeof(float) * size_t(n) + 4));
    CK(cudaMalloc(&da, sizeof(float) * size_t(n) + 4));
    CK(cudaMalloc(&dt, sizeof(float) * size_t(n) + 4));
    CK(cudaMemcpy(dx, x, sizeof(float) * size_t(n), cudaMemcpyHostToDevice));
    launch_join_math(dx, n, da, dt, nullptr);
    CK(cudaMemcpy(acos_out, da, sizeof(float) * size_t(n), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tan_out, dt, sizeof(float) * size_t(n), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(da); cudaFree(dt);
    CK(cudaGetLastError());
    return CB200_OK;
}

int64_t cb200_debug_lines(cb200_canvas *cv, float *edges, uint32_t *job_of_edge, int64_t capacity)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (cudaSetDevice(cv->device) != cudaSuccess) return CB200_ERR_CUDA;
    if (finish_pending(cv) != CB200_OK) return CB200_ERR_CUDA;
    cudaStreamSynchronize(cv->stream);
    frame_header h = *cv->pinned_hdr;
    int64_t n = int64_t(h.n_items) * 3, got = 0;
    std::vector<float4> pieces(static_cast<size_t>(n));
    std::vector<uint32_t> jobs(static_cast<size_t>(n)), rows(static_cast<size_t>(n));
    if (n) {
        cudaMemcpy(pieces.data(), cv->pieces.p, sizeof(float4) * size_t(n), cudaMemcpyDeviceToHost);
        cudaMemcpy(jobs.data(), cv->piece_job.p, sizeof(uint32_t) * size_t(n), cudaMemcpyDeviceToHost);
        cudaMemcpy(rows.data(), cv->piece_rows.p, sizeof(uint32_t) * size_t(n), cudaMemcpyDeviceToHost);
    }
    for (int64_t i = 0; i < n; ++i) {
        if (!rows[size_t(i)]) continue;
        if (got < capacity && edges) {
            memcpy(edges + got * 4, &pieces[size_t(i)], sizeof(float4));
            if (job_of_edge) job_of_edge[got] = jobs[size_t(i)] & 0x7fffffffu;      // bit 31: projected piece (raster.cu)
        }
        ++got;
    }
    return got;
}

int64_t cb200_debug_runs(cb200_canvas *cv, uint64_t *keys, float *cumulative, int64_t capacity)
{
    if (!cv) return fail(CB200_ERR_BAD_ARG, "null argument");
    if (cudaSetDevice(cv->device) != cudaSuccess) return CB200_ERR_CUDA;
    if (finish_pending(cv) != CB200_OK) return CB200_ERR_CUDA;
    cudaStreamSynchronize(cv->stream);
    frame_header h = *cv->pinned_hdr;
    int64_t n = std::min<int64_t>(h.n_runs, capacity);
    int sorted = sort_passes(cv->staged.key_bits) & 1;
    if (n && keys) cudaMemcpy(keys, sorted ? cv->keys1.p : cv->keys0.p, sizeof(uint64_t) * size_t(n), cudaMemcpyDeviceToHost);
    if (n && cumulative) cudaMemcpy(cumulative, cv->cumulative.p, sizeof(float) * size_t(n), cudaMemcpyDeviceToHost);
    return h.n_runs;
}

}  // extern "C"
