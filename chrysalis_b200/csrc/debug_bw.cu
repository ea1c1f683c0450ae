// Measurement aid (not part of the C ABI in include/): how fast can the SMs stream a [n][ld] float32
// matrix when every CTA reads a column slab of `seg` floats per row (seg * 4 bytes contiguous, rows
// ld * 4 bytes apart)?  Used by scripts/exp_segment_bw.py to pick the device layout of X.
#include "common.cuh"

namespace chr {

template <int VEC>   // float4 loads per lane per row: segment = VEC * 512 bytes
__global__ void __launch_bounds__(256) segment_read_kernel(const float* __restrict__ X, int64_t n, int64_t ld, int rows_per_cta,
                                                           float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* base = X + (int64_t)blockIdx.x * (VEC * 128) + lane * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + warp * 8; r < r1; r += 64) {
        float4 v[8][VEC];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int w = 0; w < VEC; ++w) {
                const int64_t rr = r + u < r1 ? r + u : r1 - 1;
                v[u][w] = __ldg(reinterpret_cast<const float4*>(base + rr * ld + w * 128));
            }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int w = 0; w < VEC; ++w) { acc.x += v[u][w].x; acc.y += v[u][w].y; acc.z += v[u][w].z; acc.w += v[u][w].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;   // keep the loads alive
}

}  // namespace chr

extern "C" int chr_debug_segment_read(const float* X, int64_t n, int64_t ld, int64_t ncols, int vec, int rows_per_cta, float* out,
                                      chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)(ncols / (vec * 128)), (unsigned)ceil_div<int64_t>(n, rows_per_cta));
    switch (vec) {
        case 1: segment_read_kernel<1><<<grid, 256, 0, st>>>(X, n, ld, rows_per_cta, out); break;
        case 2: segment_read_kernel<2><<<grid, 256, 0, st>>>(X, n, ld, rows_per_cta, out); break;
        case 4: segment_read_kernel<4><<<grid, 256, 0, st>>>(X, n, ld, rows_per_cta, out); break;
        default: return 2;
    }
    CHR_LAUNCH_CHECK();
    return 0;
}
