// canvas_ity.hpp -- drop-in `canvas_ity::canvas` front end for the B200 back end.
//
// Same namespace, enums, method names, argument meaning, public data members
// and silent-no-op error behaviour as a-e-k/canvas_ity v1.00 (reference
// src/canvas_ity.hpp:146-156 enums, :177-1148 class), so test/test.cpp and
// demos/tiger/tiger.cpp compile against this header unmodified.  The class only
// RECORDS: state setters and path building run on the host; every draw call is
// lowered to a cb200_draw (include/canvas_b200.h) and queued; the queue is
// flushed to the GPU at get_image_data / put_image_data / destruction.
// Defining CANVAS_ITY_IMPLEMENTATION is harmless: the implementation lives in
// libcanvas_b200.so (canvas_ity_b200/csrc).  There is no CPU rasteriser here.
#ifndef CANVAS_ITY_B200_FRONT_HPP
#define CANVAS_ITY_B200_FRONT_HPP

#include <cstddef>
#include <vector>               // the reference's interface section includes these two (hpp:139-140) ...

#ifdef CANVAS_ITY_IMPLEMENTATION
#include <algorithm>            // ... and its implementation section these (hpp:1216-1218): drivers such as
#include <cmath>                // test/test.cpp rely on getting <cmath> through CANVAS_ITY_IMPLEMENTATION
#include <numeric>
#endif

namespace canvas_ity
{

// Values are part of the contract: composite_operation is the 4-bit
// Porter-Duff mix program the compositor kernel decodes (reference :146-149).
enum composite_operation { source_in = 1, source_copy = 2, source_out = 3,
    destination_in = 4, destination_atop = 7, lighter = 10,
    destination_over = 11, destination_out = 12, source_atop = 13,
    source_over = 14, exclusive_or = 15 };
enum cap_style { butt = 0, square = 1, circle = 2 };
enum join_style { miter = 0, bevel = 1, rounded = 2 };
enum brush_type { fill_style = 0, stroke_style = 1 };
enum repetition_style { repeat = 0, repeat_x = 1, repeat_y = 2, no_repeat = 3 };
enum align_style { leftward = 0, rightward = 1, center = 2, start = 0, ending = 1 };
enum baseline_style { alphabetic = 0, top = 1, middle = 2, bottom = 3,
    hanging = 4, ideographic = 3 };

#ifndef CANVAS_ITY_B200_ENUMS_ONLY
class canvas
{
public:
    canvas( int width, int height );
    ~canvas();

    // transforms (host)
    void scale( float x, float y );
    void rotate( float angle );
    void translate( float x, float y );
    void transform( float a, float b, float c, float d, float e, float f );
    void set_transform( float a, float b, float c, float d, float e, float f );

    // compositing
    void set_global_alpha( float alpha );
    composite_operation global_composite_operation;

    // shadows
    void set_shadow_color( float red, float green, float blue, float alpha );
    float shadow_offset_x;
    float shadow_offset_y;
    void set_shadow_blur( float level );

    // line styles
    void set_line_width( float width );
    cap_style line_cap;
    join_style line_join;
    void set_miter_limit( float limit );
    float line_dash_offset;
    void set_line_dash( float const *segments, int count );

    // fill and stroke styles
    void set_color( brush_type type, float red, float green, float blue,
                    float alpha );
    void set_linear_gradient( brush_type type, float start_x, float start_y,
                              float end_x, float end_y );
    void set_radial_gradient( brush_type type, float start_x, float start_y,
                              float start_radius, float end_x, float end_y,
                              float end_radius );
    void add_color_stop( brush_type type, float offset, float red, float green,
                         float blue, float alpha );
    void set_pattern( brush_type type, unsigned char const *image, int width,
                      int height, int stride, repetition_style repetition );

    // path building (host)
    void begin_path();
    void move_to( float x, float y );
    void close_path();
    void line_to( float x, float y );
    void quadratic_curve_to( float control_x, float control_y, float x, float y );
    void bezier_curve_to( float control_1_x, float control_1_y,
                          float control_2_x, float control_2_y, float x, float y );
    void arc_to( float vertex_x, float vertex_y, float x, float y, float radius );
    void arc( float x, float y, float radius, float start_angle, float end_angle,
              bool counter_clockwise = false );
    void rectangle( float x, float y, float width, float height );

    // drawing (queued for the GPU)
    void fill();
    void stroke();
    void clip();
    bool is_point_in_path( float x, float y );
    void clear_rectangle( float x, float y, float width, float height );
    void fill_rectangle( float x, float y, float width, float height );
    void stroke_rectangle( float x, float y, float width, float height );

    // text
    align_style text_align;
    baseline_style text_baseline;
    bool set_font( unsigned char const *font, int bytes, float size );
    void fill_text( char const *text, float x, float y,
                    float maximum_width = 1.0e30f );
    void stroke_text( char const *text, float x, float y,
                      float maximum_width = 1.0e30f );
    float measure_text( char const *text );

    // images (flush points)
    void draw_image( unsigned char const *image, int width, int height,
                     int stride, float x, float y, float to_width,
                     float to_height );
    void get_image_data( unsigned char *image, int width, int height,
                         int stride, int x, int y );
    void put_image_data( unsigned char const *image, int width, int height,
                         int stride, int x, int y );

    // state stack
    void save();
    void restore();

    // ---- extensions beyond the reference API (all optional) ----
    struct host_state;
    host_state *b200() { return self; }     // back-end access for bindings/tests
    canvas( int width, int height, int device, int band_y0, int band_rows );

private:
    host_state *self;
    canvas( canvas const & );
    canvas &operator=( canvas const & );
};
#endif // CANVAS_ITY_B200_ENUMS_ONLY

}

#endif // CANVAS_ITY_B200_FRONT_HPP
