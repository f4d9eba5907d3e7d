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

void launch_fill_f32(float *dst, float value, uint64_t n, cudaStream_t s)
{
    if (!n) return;
    launch_pdl(k_fill, grid_for(n), kBlock, 0, s, dst, value, n);
}

}  // namespace cb200
