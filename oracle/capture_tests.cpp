// capture_tests.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Runs every entry of the reference's own test table (test/test.cpp:2185-2262,
// compiled from /root/reference via TEST_CPP, a stripped temp copy that only
// lacks the `#include "../src/canvas_ity.hpp"` line) against the capture shim and
// writes, per test: <name>.cvs (API call stream), <name>.rgba8 (the reference's
// get_image_data output), <name>.f32 (its float framebuffer) and manifest.txt
// (name, expected hash, size).  tools/make_golden.py packs these into
// tests/golden/.
#define canvas_ity canvas_ity_ref
#define CANVAS_ITY_IMPLEMENTATION
#include REFERENCE_HPP
#undef canvas_ity

#include "capture_shim.hpp"

#define main reference_test_main
#include TEST_CPP
#undef main

#include <cstdio>

static void write_file(std::string const &path, void const *data, size_t bytes)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(1); }
    fwrite(data, 1, bytes, f);
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: capture_tests <outdir>\n"); return 2; }
    std::string out = argv[1];
    base64_decode(font_a_base64, font_a);
    base64_decode(font_b_base64, font_b);
    base64_decode(font_c_base64, font_c);
    base64_decode(font_d_base64, font_d);
    base64_decode(font_e_base64, font_e);
    base64_decode(font_f_base64, font_f);
    base64_decode(font_g_base64, font_g);
    FILE *manifest = fopen((out + "/manifest.txt").c_str(), "w");
    size_t count = sizeof(tests) / sizeof(tests[0]);
    for (size_t i = 0; i < count; ++i) {
        test const &t = tests[i];
        canvas_ity::canvas subject(t.width, t.height);
        t.call(subject, static_cast<float>(t.width), static_cast<float>(t.height));
        subject.sync();
        std::vector<unsigned char> image(size_t(4 * t.width * t.height));
        subject.real.get_image_data(&image[0], t.width, t.height, 4 * t.width, 0, 0);
        unsigned hash = hash_image(&image[0], t.width, t.height);
        std::string base = out + "/" + t.name;
        write_file(base + ".cvs", subject.script.bytes.data(), subject.script.bytes.size());
        write_file(base + ".rgba8", &image[0], image.size());
        write_file(base + ".f32", subject.real.bitmap, sizeof(float) * 4 * size_t(t.width * t.height));
        fprintf(manifest, "%s %08x %08x %d %d\n", t.name, t.hash, hash, t.width, t.height);
    }
    fclose(manifest);
    return 0;
}
