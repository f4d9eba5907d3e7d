// capture_tiger.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Compiles the reference's tiger demo (demos/tiger/tiger.cpp:96-4355, read from
// /root/reference through TIGER_CPP, a stripped temp copy without its
// `#include "../../src/canvas_ity.hpp"` line) against the capture shim and writes
// the call stream of its first frame -- 305 draws, 2222 cubics -- as a canvas
// script.  The demo's own readback is not recorded.
#define canvas_ity canvas_ity_ref
#define CANVAS_ITY_IMPLEMENTATION
#include REFERENCE_HPP
#undef canvas_ity

#define CAPTURE_SKIP_GET_IMAGE_DATA
#include "capture_shim.hpp"

#include <cstdio>
#include <cstdlib>

static char const *g_out = 0;

static void first_frame_done(canvas_ity::canvas &c)
{
    c.sync();
    FILE *f = fopen(g_out, "wb");
    if (!f) { perror(g_out); exit(1); }
    fwrite(c.script.bytes.data(), 1, c.script.bytes.size(), f);
    fclose(f);
    exit(0);
}

#define main tiger_demo_main
#include TIGER_CPP
#undef main

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: capture_tiger <out.cvs>\n"); return 2; }
    g_out = argv[1];
    canvas_ity::canvas::on_destroy = first_frame_done;
    return tiger_demo_main();
}
