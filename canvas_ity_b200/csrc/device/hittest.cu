// hittest.cu -- bulk is_point_in_path (reference hpp:3101-3132) for many query points at once
// (SURVEY 8f-3).  The reference answers one point per call by walking every edge of the flattened
// path on the CPU; here the path is flattened once (host, the same routine K1 uses) and every query
// point is one thread:
//
//   k_hit_test     grid (query blocks, edge chunks): a CTA stages 256 edges at a time in shared
//                  memory (one coalesced 16 B load per thread, then conflict-free broadcasts) and
//                  each thread runs the reference's crossing rule against them -- signed crossings
//                  of the half-open span (from.y, to.y], "on the edge" when the side product is
//                  exactly zero or the point lies on a horizontal edge.  The side product is the
//                  reference's dot(perpendicular(to - from), point - from), unfused (-fmad=false), so
//                  the exact-zero test and the sign agree bit for bit.
//   k_hit_resolve  inside = on an edge || winding != 0 (the early `return true` of the reference is
//                  order independent, so edge chunks simply add up).
// Compute bound: n_queries x n_edges rule evaluations, ~14 instructions each.
#include "frame.cuh"

#include <algorithm>

namespace cb200 {

namespace {

constexpr int kHitBlock = 256;

__global__ void __launch_bounds__(kHitBlock) k_hit_test(const float4 *edges, uint32_t n_edges, const float2 *queries,
                                                        uint32_t n_queries, int2 *acc, uint32_t edges_per_chunk)
{
    __shared__ float4 tile[kHitBlock];
    __shared__ float2 span[kHitBlock];                      // (min y, max y) of the staged edges
    const uint32_t q = blockIdx.x * kHitBlock + threadIdx.x;
    const bool live = q < n_queries;
    const float2 at = live ? queries[q] : make_float2(0.0f, 0.0f);
    const vec2 p = v2(at.x, at.y);
    const uint32_t e0 = blockIdx.y * edges_per_chunk, e1 = min(n_edges, e0 + edges_per_chunk);
    int winding = 0, on_edge = 0;
    for (uint32_t base = e0; base < e1; base += kHitBlock) {
        __syncthreads();
        if (base + threadIdx.x < e1) {
            const float4 e = edges[base + threadIdx.x];
            tile[threadIdx.x] = e;
            // a NaN ordinate fails every comparison of the reference: give such an edge an empty span
            const bool ordered = e.y == e.y && e.w == e.w;
            span[threadIdx.x] = ordered ? make_float2(fminf(e.y, e.w), fmaxf(e.y, e.w)) : make_float2(1.0f, 0.0f);
        }
        __syncthreads();
        const int n = int(min(uint32_t(kHitBlock), e1 - base));
        // The reference's (from.y < y && y <= to.y) || (to.y < y && y <= from.y) is lo < y && y <= hi;
        // its horizontal-edge rule needs lo == y == hi.  Both live in lo <= y && y <= hi, which most
        // edges fail, so the common path is one 8-byte broadcast load and two compares.
#pragma unroll 8
        for (int k = 0; k < n; ++k) {
            const float2 sp = span[k];
            if (sp.x <= p.y && p.y <= sp.y) {
                const float4 e = tile[k];
                const vec2 from = v2(e.x, e.y), to = v2(e.z, e.w);
                if (sp.x < p.y) {
                    const float side = dot(perp(to - from), p - from);
                    if (side == 0.0f) on_edge = 1;
                    else winding += side > 0.0f ? 1 : -1;
                } else if (sp.x == sp.y && ((from.x <= p.x && p.x <= to.x) || (to.x <= p.x && p.x <= from.x)))
                    on_edge = 1;
            }
        }
    }
    if (live) {
        if (winding) atomicAdd(&acc[q].x, winding);
        if (on_edge) atomicOr(&acc[q].y, 1);
    }
}

__global__ void __launch_bounds__(kHitBlock) k_hit_resolve(const int2 *acc, uint32_t n_queries, uint8_t *inside)
{
    const uint32_t q = blockIdx.x * kHitBlock + threadIdx.x;
    if (q < n_queries) { const int2 a = acc[q]; inside[q] = (a.y || a.x) ? 1 : 0; }
}

}  // namespace

void launch_hit_test(const float4 *edges, uint32_t n_edges, const float2 *queries, uint32_t n_queries, int2 *acc,
                     uint8_t *inside, cudaStream_t s)
{
    if (!n_queries) return;
    const uint32_t q_blocks = (n_queries + kHitBlock - 1) / kHitBlock;
    // few queries against a long path: split the edges too, so that the grid still fills the GPU
    uint32_t chunks = 1;
    const uint32_t edge_tiles = (n_edges + kHitBlock - 1) / kHitBlock;
    if (q_blocks < 4u * kSMs && edge_tiles > 1) chunks = std::min(edge_tiles, (4u * kSMs + q_blocks - 1) / q_blocks);
    const uint32_t per_chunk = std::max(1u, (edge_tiles + chunks - 1) / chunks) * kHitBlock;
    chunks = n_edges ? (n_edges + per_chunk - 1) / per_chunk : 1;
    cudaMemsetAsync(acc, 0, sizeof(int2) * size_t(n_queries), s);
    k_hit_test<<<dim3(q_blocks, chunks), kHitBlock, 0, s>>>(edges, n_edges, queries, n_queries, acc, per_chunk);
    k_hit_resolve<<<q_blocks, kHitBlock, 0, s>>>(acc, n_queries, inside);
}

}  // namespace cb200
