// ref_png.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The reference test driver's own PNG writer (write_png, test/test.cpp:2415-2507), compiled from
// where it lies (TEST_CPP = a temp copy of /root/reference/test/test.cpp that only lacks its
// `#include "../src/canvas_ity.hpp"` line; nothing is copied into this repo) and exported as
// ref_write_png().  It is the byte-exact golden for cb200_encode_png.  The driver needs the library
// it tests, so the reference header comes along under a private namespace (the flat API in
// ref_api.cpp holds the real one).
#define canvas_ity canvas_ity_png_driver
#define CANVAS_ITY_IMPLEMENTATION
#include REFERENCE_HPP

#define main reference_test_main
#include TEST_CPP
#undef main
#undef canvas_ity

extern "C" void ref_write_png(const char *path, const unsigned char *rgba, int width, int height)
{
    write_png(path, rgba, width, height);
}
