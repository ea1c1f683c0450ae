// Shared definitions of the square-lattice stencil kernels: moran_lattice.cu (Moran's I / Geary's C) and moran_perm.cu
// (permutation inference).  Layout, band / chunk tables and the walking scheme are described in moran_lattice.cu.
#pragma once
#include "async.cuh"
#include "moran.cuh"

namespace chr {

constexpr int kLatR = 6;                                          // lattice rows per compute warp
constexpr int kLatWarps = 10;                                     // compute warps
#ifndef LAT_PRODUCERS
#define LAT_PRODUCERS 4
#endif
constexpr int kLatProducers = LAT_PRODUCERS;                      // producer warps
constexpr int kLatBH = kLatR * kLatWarps;                         // 60 lattice rows per band
constexpr int kLatSegRows = kLatBH + 1;                           // + the row below the band
constexpr int kLatRowBytes = kGeneTile * 4;                       // 512
constexpr int kLatSegBytes = kLatSegRows * kLatRowBytes;          // one staged column
constexpr int kLatStageCols = 2;
constexpr int kLatStages = 3;
constexpr int kLatStageBytes = kLatStageCols * kLatSegBytes;
constexpr int kLatThreads = 32 * (kLatWarps + kLatProducers);
constexpr int kLatFlushStages = 4;                                // level-1 sums cover 4 stages = 8 columns
constexpr size_t kLatSmem = 128 + (size_t)kLatStages * kLatStageBytes + 16 * kLatStages;
constexpr int kLatPilotMax = 512;
constexpr int kCorrPerCta = 256;                                  // correction entries per CTA (8 batches of 4 per warp)
constexpr int kCorrWarps = 8;

static_assert(kLatBH == CHR_LATTICE_BAND_ROWS, "header and kernel disagree on the band height");

__device__ __forceinline__ Row4 lat_lds_row4(uint32_t addr) {
    Row4 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.a), "=l"(r.b) : "r"(addr));
    return r;
}
__device__ __forceinline__ void lat_cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void lat_cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void lat_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a lost completion becomes a trap (an error code), never a hung GPU
__device__ __forceinline__ void lat_mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

}  // namespace chr
