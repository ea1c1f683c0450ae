// Shared helpers for the chrysalis_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/chrysalis_b200.h"

namespace chr {

void set_error(const char* fmt, ...);

// two events around the dominant kernel of a family (CHR_FLAG_TIMING)
struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool armed = false, started = false;
};
KernelTimer& timer(int family);
void timer_begin(int family, bool on, cudaStream_t s);
void timer_end(int family, bool on, cudaStream_t s);
bool timing_env();   // CHR_TIMING=1: entry points without a flags argument (kNN, plan, prep, AA) record their kernel times too

constexpr int kNumSMs = 148;  // B200

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace chr

#define CHR_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            chr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define CHR_REQUIRE(cond, ...)           \
    do {                                 \
        if (!(cond)) {                   \
            chr::set_error(__VA_ARGS__); \
            return 2;                    \
        }                                \
    } while (0)

#define CHR_LAUNCH_CHECK() CHR_CUDA(cudaGetLastError())
