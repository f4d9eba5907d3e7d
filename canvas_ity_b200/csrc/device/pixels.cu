// pixels.cu -- K10 readback (get_image_data, hpp:3348-3381), K11 upload
// (put_image_data hpp:3383-3408 and the texel conversion of set_pattern
// hpp:2852-2860).  Pure streaming kernels: 16 B in / 4 B out per pixel (or the
// reverse), one thread per pixel, consecutive threads on consecutive pixels.
#include "frame.cuh"

#include <algorithm>

namespace cb200 {

namespace {

// delinearized, hpp:1282-1284.  __powf (ex2(y * lg2(x)) on the SFU) is good to ~1e-6 relative,
// i.e. 3e-4 of an 8-bit step: the kernel turns from ALU-bound (135 M instructions for 4096^2 with
// powf) into a memory-bound one, and stays within the +-1 LSB readback tolerance.
__device__ __forceinline__ float to_srgb(float v)
{
    return v < 0.0031308f ? 12.92f * v : 1.055f * __powf(v, 1.0f / 2.4f) - 0.055f;
}

__device__ __forceinline__ float to_linear(float v)   // linearized, hpp:1273-1275
{
    return v < 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
}

// 4x4 ordered-dither thresholds (k + 0.5) / 16 of the classic Bayer matrix, hpp:3358-3362
__constant__ float c_bayer[16] = {
    0.5f / 16, 8.5f / 16, 2.5f / 16, 10.5f / 16, 12.5f / 16, 4.5f / 16, 14.5f / 16, 6.5f / 16,
    3.5f / 16, 11.5f / 16, 1.5f / 16, 9.5f / 16, 15.5f / 16, 7.5f / 16, 13.5f / 16, 5.5f / 16 };

// One pixel of get_image_data (hpp:3364-3379): unpremultiply, clamp, delinearise, ordered dither.
__device__ __forceinline__ uchar4 encode(float4 c, int cx, int cy, bool bgra)
{
    if (c.w < kThreshold) c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);          // unpremultiplied()
    else { float k = 1.0f / c.w; c.x = k * c.x; c.y = k * c.y; c.z = k * c.z; }
    float th = c_bayer[(cy & 3) * 4 + (cx & 3)];
    uchar4 o;
    o.x = static_cast<unsigned char>(th + 255.0f * to_srgb(clamp01(c.x)));
    o.y = static_cast<unsigned char>(th + 255.0f * to_srgb(clamp01(c.y)));
    o.z = static_cast<unsigned char>(th + 255.0f * to_srgb(clamp01(c.z)));
    o.w = static_cast<unsigned char>(th + 255.0f * clamp01(c.w));
    if (bgra) { unsigned char t = o.x; o.x = o.z; o.z = t; }               // TGA / BMP channel order
    return o;
}

// grid: (column chunks, row stride).  A thread converts kReadbackPixels pixels of one row that are
// blockDim.x apart, so their four 16-byte loads are in flight together and every warp access
// stays contiguous; no 64-bit division anywhere.
constexpr int kReadbackPixels = 4;

__global__ void __launch_bounds__(kBlock) k_readback(const float4 *fb, int width, int band_y0, int band_rows,
                                                      uchar4 *dst, int dst_w, int dst_h, int ox, int oy, int bgra)
{
    grid_dependency_wait();
    const int span = kBlock * kReadbackPixels;
    for (int iy = blockIdx.y; iy < dst_h; iy += gridDim.y) {
        const int cy = oy + iy;
        const bool row_ok = cy >= band_y0 && cy < band_y0 + band_rows;
        const float4 *src = fb + size_t(row_ok ? cy - band_y0 : 0) * size_t(width);
        uchar4 *out = dst + size_t(iy) * size_t(dst_w);
        for (int base = blockIdx.x * span; base < dst_w; base += gridDim.x * span) {
            float4 c[kReadbackPixels];
#pragma unroll
            for (int k = 0; k < kReadbackPixels; ++k) {
                const int ix = base + k * kBlock + int(threadIdx.x), cx = ox + ix;
                c[k] = (ix < dst_w && row_ok && cx >= 0 && cx < width) ? __ldcs(src + cx) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int k = 0; k < kReadbackPixels; ++k) {
                const int ix = base + k * kBlock + int(threadIdx.x);
                if (ix < dst_w) out[ix] = encode(c[k], ox + ix, cy, bgra != 0);
            }
        }
    }
}

__device__ __forceinline__ float4 decode(uchar4 t)
{
    float a = t.w / 255.0f;
    return make_float4(to_linear(t.x / 255.0f) * a, to_linear(t.y / 255.0f) * a,
                       to_linear(t.z / 255.0f) * a, a);
}

__global__ void __launch_bounds__(kBlock) k_upload(float4 *fb, int width, int band_y0, int band_rows,
                                                    const uchar4 *src, int src_w, int src_h, int ox, int oy)
{
    grid_dependency_wait();
    size_t n = size_t(src_w) * size_t(src_h);
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        int cx = ox + int(i % size_t(src_w)), cy = oy + int(i / size_t(src_w));
        if (cx < 0 || cx >= width || cy < band_y0 || cy >= band_y0 + band_rows) continue;
        fb[size_t(cy - band_y0) * size_t(width) + size_t(cx)] = decode(src[i]);
    }
}

__global__ void __launch_bounds__(kBlock) k_texels(const uchar4 *src, float4 *dst, uint64_t n)
{
    grid_dependency_wait();
    uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = decode(src[i]);
}

__global__ void __launch_bounds__(kBlock) k_fill(float *dst, float value, uint64_t n)
{
    grid_dependency_wait();
    uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

// ---- PNG ------------------------------------------------------------------------------------------
// The file the reference's test driver writes after get_image_data (write_png, test/test.cpp:2415-2507):
// signature, IHDR, sRGB, one IDAT holding a zlib stream of STORED deflate blocks -- one block per row:
// 5 header bytes, filter byte 0, the row's RGBA8 -- closed by Adler-32 and the chunk CRC-32, then IEND.
// The reference does this on the CPU, two table look-ups and two modulo operations per byte after the
// readback.  Here ONE kernel reads the float framebuffer, does the sRGB + dither conversion of
// get_image_data (hpp:3348-3381), writes the bytes where they belong in the file and folds both
// checksums in on the way:
//   * CRC-32 is linear over GF(2): crc(A || B) = crc(A) * x^(8|B|) mod P  xor  crc(B).  A lane takes the
//     raw CRC of its own 32 contiguous bytes (slice-by-4 tables in shared memory), multiplies it by
//     x^(bytes that follow it in its row) (per-column constants), the warp xors the lanes together,
//     multiplies by x^(bytes that follow the row) (per-row constants) and keeps a running xor.
//   * Adler-32: A = 1 + sum b_i, B = N + sum (N - i) b_i (mod 65521), i the position in the uncompressed
//     stream -- plain sums, taken four bytes at a time with dp4a.
// Integer work throughout: the file is byte-identical to the reference's for the same pixels.
constexpr uint32_t kCrcPoly = 0xedb88320u;
constexpr uint32_t kAdlerMod = 65521u;
constexpr int kPngSeg = 256;                  // pixels per warp step: 8 consecutive pixels per lane
constexpr int kPngBlock = 256;

// a * b mod P in the reflected representation (bit 31 = x^0)
__device__ inline uint32_t gf_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
    }
    return p;
}

__device__ inline uint32_t gf_xpow(uint64_t n)          // x^n mod P
{
    uint32_t sq = 0x40000000u, p = 0x80000000u;
    while (n) {
        if (n & 1u) p = gf_mul(sq, p);
        sq = gf_mul(sq, sq);
        n >>= 1;
    }
    return p;
}

__device__ inline uint32_t crc_byte(uint32_t s, uint32_t byte)       // bitwise, for the handful of framing bytes
{
    s ^= byte;
    for (int k = 0; k < 8; ++k) s = (s >> 1) ^ ((s & 1u) ? kCrcPoly : 0u);
    return s;
}

// tables: [0, 1024) slice-by-4 CRC tables (T_k[i] = byte i followed by k zero bytes), then
// col_shift[0 .. width] = x^(32 j) (j pixels follow in the row), then row_shift[0 .. height) =
// x^(8 ((height - 1 - y) row_len + 4)) (the rows below and the Adler-32 follow row y).
__global__ void __launch_bounds__(kPngBlock) k_png_tables(uint32_t *tables, int width, int height)
{
    const uint32_t row_len = 6u + 4u * uint32_t(width);
    const int n = max(max(width + 1, height), 256);
    for (int i = blockIdx.x * kPngBlock + threadIdx.x; i < n; i += gridDim.x * kPngBlock) {
        if (i < 256) {
            uint32_t t = crc_byte(0u, uint32_t(i));
            for (int k = 0; k < 4; ++k) { tables[k * 256 + i] = t; t = crc_byte(t, 0u); }
        }
        if (i <= width) tables[1024 + i] = gf_xpow(32ull * uint64_t(i));
        if (i < height) tables[1024 + width + 1 + i] = gf_xpow(8ull * (uint64_t(height - 1 - i) * row_len + 4ull));
    }
}

__global__ void __launch_bounds__(kPngBlock, 3) k_png_rows(const float4 *fb, int width, int height, uint8_t *out,
                                                         const uint32_t *tables, png_sums *acc, uint32_t *row_crc)
{
    __shared__ uint32_t T[1024];
    __shared__ uint32_t staged[kPngBlock / 32][kPngSeg + kPngSeg / 8];     // word i at i + i / 8
    for (int i = threadIdx.x; i < 1024; i += kPngBlock) T[i] = tables[i];
    __syncthreads();
    const uint32_t *col_shift = tables + 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *mine = staged[warp];
    const int segs_per_row = (width + kPngSeg - 1) / kPngSeg;
    const long long n_segs = (long long)segs_per_row * height;
    const uint32_t row_bytes = 1u + 4u * uint32_t(width), row_len = 5u + row_bytes;
    const uint64_t n_total = uint64_t(height) * row_bytes;               // uncompressed length
    unsigned long long sum_a = 0, sum_b = 0;
    for (long long sidx = (long long)blockIdx.x * (kPngBlock / 32) + warp; sidx < n_segs;
         sidx += (long long)gridDim.x * (kPngBlock / 32)) {
        const int y = int(sidx / segs_per_row), x0 = int(sidx % segs_per_row) * kPngSeg;
        const int n_valid = min(kPngSeg, width - x0);
        const float4 *src = fb + size_t(y) * size_t(width) + size_t(x0);
        float4 c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ix = k * 32 + lane;
            c[k] = ix < n_valid ? __ldcs(src + ix) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ix = k * 32 + lane;
            const uchar4 px = encode(c[k], x0 + ix, y, false);
            mine[ix + (ix >> 3)] = ix < n_valid ? (uint32_t(px.x) | uint32_t(px.y) << 8 | uint32_t(px.z) << 16 | uint32_t(px.w) << 24) : 0u;
        }
        __syncwarp();
        // checksums over this lane's own run of up to 8 consecutive pixels
        const int run0 = 8 * lane, run_n = max(0, min(8, n_valid - run0));
        const uint64_t first_byte = uint64_t(y) * row_bytes + 1u + 4ull * uint64_t(x0 + run0);
        uint32_t weight = uint32_t((n_total - first_byte) % kAdlerMod);     // (N - i) mod 65521 of the run's first byte
        uint32_t crc = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < run_n) {
                const uint32_t w = mine[9 * lane + k];
                crc ^= w;
                crc = T[768 + (crc & 0xffu)] ^ T[512 + ((crc >> 8) & 0xffu)] ^ T[256 + ((crc >> 16) & 0xffu)] ^ T[crc >> 24];
                const uint32_t s4 = __dp4a(w, 0x01010101u, 0u), off = __dp4a(w, 0x03020100u, 0u);
                sum_a += s4;
                sum_b += (unsigned long long)weight * s4 + (2u * kAdlerMod - off);      // + 2p keeps it non-negative
                weight = weight >= 4u ? weight - 4u : weight + kAdlerMod - 4u;
            }
        }
        uint32_t part = run_n ? gf_mul(col_shift[width - (x0 + run0 + run_n)], crc) : 0u;
        uint8_t *row_out = out + 56 + size_t(y) * size_t(row_len);
        if (x0 == 0 && lane == 0) {
            // stored-block header + filter byte (test.cpp:2472-2481), ahead of the row's pixels
            const uint32_t rs = row_bytes;
            const uint32_t b[6] = { uint32_t(y + 1 == height), rs & 255u, (rs >> 8) & 255u, ~rs & 255u, (~rs >> 8) & 255u, 0u };
            uint32_t ps = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) ps = T[(ps ^ b[k]) & 0xffu] ^ (ps >> 8);
            part ^= gf_mul(col_shift[width], ps);
            uint16_t *h16 = reinterpret_cast<uint16_t *>(row_out);             // 56 + y * row_len is even
            h16[0] = uint16_t(b[0] | b[1] << 8); h16[1] = uint16_t(b[2] | b[3] << 8); h16[2] = uint16_t(b[4] | b[5] << 8);
        }
        // the row's CRC (relative to the row's end) collects in row_crc[y]; k_png_finish shifts the rows
        const uint32_t seg_crc = __reduce_xor_sync(0xffffffffu, part);
        if (lane == 0 && seg_crc) atomicXor(&row_crc[y], seg_crc);
        // the pixels: rows start 2 bytes past a word boundary when y is even, so words are re-cut there
        uint8_t *seg_out = row_out + 6 + 4 * size_t(x0);
        if ((reinterpret_cast<uintptr_t>(seg_out) & 3u) == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int ix = k * 32 + lane;
                if (ix < n_valid) reinterpret_cast<uint32_t *>(seg_out)[ix] = mine[ix + (ix >> 3)];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int j = k * 32 + lane;                                    // output word j straddles pixels j - 1 and j
                if (j > n_valid) continue;
                const uint32_t lo = j > 0 ? mine[(j - 1) + ((j - 1) >> 3)] : 0u, hi = j < n_valid ? mine[j + (j >> 3)] : 0u;
                const uint32_t v = __funnelshift_r(lo, hi, 16);
                uint8_t *at = seg_out - 2 + 4 * size_t(j);
                if (j == 0) *reinterpret_cast<uint16_t *>(at + 2) = uint16_t(v >> 16);
                else if (j == n_valid) *reinterpret_cast<uint16_t *>(at) = uint16_t(v);
                else *reinterpret_cast<uint32_t *>(at) = v;
            }
        }
        __syncwarp();
    }
    const uint32_t ra = __reduce_add_sync(0xffffffffu, uint32_t(sum_a % kAdlerMod));
    const uint32_t rb = __reduce_add_sync(0xffffffffu, uint32_t(sum_b % kAdlerMod));
    if (lane == 0) {
        atomicAdd(&acc->a, (unsigned long long)ra);
        atomicAdd(&acc->b, (unsigned long long)rb);
    }
}

// Header, Adler-32, IDAT CRC and IEND (test.cpp:2430-2459, 2491-2506).  One CTA: the threads shift every
// row's CRC past the rows below it (x^(8 ((height - 1 - y) row_len + 4)), k_png_tables) and xor them
// together, thread 0 writes the framing.
__global__ void __launch_bounds__(1024) k_png_finish(uint8_t *out, int width, int height, const png_sums *acc,
                                                     const uint32_t *tables, const uint32_t *row_crc)
{
    __shared__ uint32_t warp_xor[32];
    const uint32_t *row_shift = tables + 1024 + (width + 1);
    uint32_t rows_crc = 0;
    for (int y = threadIdx.x; y < height; y += blockDim.x)
        if (const uint32_t c = row_crc[y]) rows_crc ^= gf_mul(row_shift[y], c);
    rows_crc = __reduce_xor_sync(0xffffffffu, rows_crc);
    if ((threadIdx.x & 31) == 0) warp_xor[threadIdx.x >> 5] = rows_crc;
    __syncthreads();
    if (threadIdx.x) return;
    rows_crc = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) rows_crc ^= warp_xor[w];
    const uint32_t row_bytes = 1u + 4u * uint32_t(width), row_len = 5u + row_bytes;
    const uint32_t idat_size = 6u + uint32_t(height) * row_len;
    const uint64_t n_total = uint64_t(height) * row_bytes;
    uint8_t header[56] = { 137, 80, 78, 71, 13, 10, 26, 10, 0, 0, 0, 13, 73, 72, 68, 82,
                           uint8_t(width >> 24), uint8_t(width >> 16), uint8_t(width >> 8), uint8_t(width),
                           uint8_t(height >> 24), uint8_t(height >> 16), uint8_t(height >> 8), uint8_t(height),
                           8, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 115, 82, 71, 66, 0, 174, 206, 28, 233,
                           uint8_t(idat_size >> 24), uint8_t(idat_size >> 16), uint8_t(idat_size >> 8), uint8_t(idat_size),
                           73, 68, 65, 84, 120, 1 };
    uint32_t crc = ~0u;
    for (int i = 12; i < 29; ++i) crc = crc_byte(crc, header[i]);
    header[29] = uint8_t(~crc >> 24); header[30] = uint8_t(~crc >> 16); header[31] = uint8_t(~crc >> 8); header[32] = uint8_t(~crc);
    for (int i = 0; i < 56; ++i) out[i] = header[i];
    const uint32_t a = uint32_t((1ull + acc->a) % kAdlerMod), b = uint32_t((n_total % kAdlerMod + acc->b) % kAdlerMod);
    uint8_t footer[20] = { uint8_t(b >> 8), uint8_t(b), uint8_t(a >> 8), uint8_t(a), 0, 0, 0, 0, 0, 0, 0, 0, 73, 69, 78, 68, 174, 66, 96, 130 };
    // IDAT CRC: "IDAT" + zlib header from the all-ones start, shifted past everything that follows,
    // xor the rows' combined raw CRC (already shifted past the Adler bytes), xor the Adler bytes' own
    uint32_t head = ~0u;
    for (int i = 50; i < 56; ++i) head = crc_byte(head, header[i]);
    uint32_t tail = 0;
    for (int i = 0; i < 4; ++i) tail = crc_byte(tail, footer[i]);
    const uint32_t state = gf_mul(gf_xpow(8ull * (uint64_t(height) * row_len + 4ull)), head) ^ rows_crc ^ tail;
    footer[4] = uint8_t(~state >> 24); footer[5] = uint8_t(~state >> 16); footer[6] = uint8_t(~state >> 8); footer[7] = uint8_t(~state);
    uint8_t *end = out + 56 + size_t(height) * size_t(row_len);
    for (int i = 0; i < 20; ++i) end[i] = footer[i];
}

inline int grid_for(uint64_t n)
{
    uint64_t blocks = (n + kBlock - 1) / kBlock;
    uint64_t cap = uint64_t(148) * 16;
    return int(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

void launch_readback(const float4 *fb, int width, int band_y0, int band_rows, uint8_t *dst, int dst_w,
                     int dst_h, int x, int y, cudaStream_t s, int bgra)
{
    if (dst_w <= 0 || dst_h <= 0) return;
    const int span = kBlock * kReadbackPixels;
    const int chunks = std::min((dst_w + span - 1) / span, 16);
    const int rows = std::min(dst_h, std::max(1, (148 * 16) / chunks));
    launch_pdl(k_readback, dim3(chunks, rows), kBlock, 0, s, fb, width, band_y0, band_rows, reinterpret_cast<uchar4 *>(dst),
               dst_w, dst_h, x, y, bgra);
}

void launch_upload(float4 *fb, int width, int band_y0, int band_rows, const uint8_t *src, int src_w,
                   int src_h, int x, int y, cudaStream_t s)
{
    uint64_t n = uint64_t(src_w) * uint64_t(src_h);
    if (!n) return;
    launch_pdl(k_upload, grid_for(n), kBlock, 0, s, fb, width, band_y0, band_rows, reinterpret_cast<const uchar4 *>(src),
                                            src_w, src_h, x, y);
}

void launch_texel_convert(const uint8_t *src, float4 *dst, uint64_t n_texels, cudaStream_t s)
{
    if (!n_texels) return;
    launch_pdl(k_texels, grid_for(n_texels), kBlock, 0, s, reinterpret_cast<const uchar4 *>(src), dst, n_texels);
}

size_t png_table_words(int width, int height) { return size_t(1024) + size_t(width) + 1 + size_t(height); }

void launch_png_tables(uint32_t *tables, int width, int height, cudaStream_t s)
{
    const int n = std::max(std::max(width + 1, height), 256);
    k_png_tables<<<(n + kPngBlock - 1) / kPngBlock, kPngBlock, 0, s>>>(tables, width, height);
}

void launch_png_encode(const float4 *fb, int width, int height, uint8_t *out, const uint32_t *tables, png_sums *acc,
                       uint32_t *row_crc, cudaStream_t s)
{
    cudaMemsetAsync(acc, 0, sizeof(png_sums), s);
    cudaMemsetAsync(row_crc, 0, sizeof(uint32_t) * size_t(height), s);
    const long long segs = (long long)((width + kPngSeg - 1) / kPngSeg) * height;
    const int ctas = int(std::min<long long>((segs + kPngBlock / 32 - 1) / (kPngBlock / 32), 4 * kSMs));
    k_png_rows<<<std::max(ctas, 1), kPngBlock, 0, s>>>(fb, width, height, out, tables, acc, row_crc);
    k_png_finish<<<1, 1024, 0, s>>>(out, width, height, acc, tables, row_crc);
}

void launch_fill_f32(float *dst, float value, uint64_t n, cudaStream_t s)
{
    if (!n) return;
    launch_pdl(k_fill, grid_for(n), kBlock, 0, s, dst, value, n);
}

}  // namespace cb200
