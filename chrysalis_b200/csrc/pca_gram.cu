// K6 - PCA covariance (Gram) of the SVG matrix on tcgen05 tensor cores.
//
// Replaces the factorisation inside  PCA(n_components, svd_solver='arpack').fit_transform(pcs)
// (/root/reference/chrysalis/core.py:126-128; sklearn _pca.py centres X and runs ARPACK svds).
// Here:  C = Z^T Z  with Z = P - mean  (the only dense contraction on the path), then a small
// eigensolver (pca_eig.cu) and the scores GEMM (pca_scores.cu).
//
// Precision: operands are split bf16 pairs  z = hi + lo  (16 mantissa bits) and the tile
// accumulates  hi_i hi_j^T + hi_i lo_j^T + lo_i hi_j^T  in fp32 TMEM (error ~2^-16 per product,
// random), over at most 8192 spots per accumulation chain; the split-K partial tiles are
// summed in fp64.  A single tf32/bf16 pass is not enough for the 0.9999 subspace bound when
// eigen-gaps are ~1e-3.
//
// Data flow:  P [n][ld] fp32  --centre/split/transpose-->  Zt_hi, Zt_lo [s_pad][n_pad] bf16
// (K = spots contiguous: both MMA operands are K-major, 128-byte swizzled TMA tiles)
//   --tcgen05 (upper-triangular 128x128 tiles x K splits)-->  fp32 partial tiles
//   --fp64 reduce + mirror-->  C [s][s] fp64.
#include <cuda_bf16.h>

#include "async.cuh"

namespace chr {

// ---------------------------------------------------------------------------------------------
// column sums (fp64), two stages, fixed order
// ---------------------------------------------------------------------------------------------
constexpr int kCsRows = 4096;  // spots per partial

__global__ void colsum_partial_kernel(const float* __restrict__ P, int64_t n, int64_t s, int64_t ld,
                                      double* __restrict__ part) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= s) return;
    const int64_t r0 = (int64_t)blockIdx.y * kCsRows, r1 = r0 + kCsRows < n ? r0 + kCsRows : n;
    double acc = 0.0;
    for (int64_t r = r0; r < r1; ++r) acc += (double)__ldg(P + r * ld + col);
    part[(int64_t)blockIdx.y * s + col] = acc;
}

__global__ void colsum_final_kernel(const double* __restrict__ part, int nsplit, int64_t s, double* __restrict__ sums) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= s) return;
    double a = 0.0;
    for (int p = 0; p < nsplit; ++p) a += part[(int64_t)p * s + col];
    sums[col] = a;
}

// ---------------------------------------------------------------------------------------------
// centre, split into bf16 hi/lo, transpose:  Zt[g][i] = P[i][g] - mean[g]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
center_split_transpose_kernel(const float* __restrict__ P, int64_t n, int64_t s, int64_t ld, const double* __restrict__ mean,
                              __nv_bfloat16* __restrict__ Zhi, __nv_bfloat16* __restrict__ Zlo, int64_t n_pad) {
    __shared__ float tile[64][65];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    const int64_t g0 = (int64_t)blockIdx.x * 64, i0 = (int64_t)blockIdx.y * 64;
    const int64_t g = g0 + tx;
    const float mu = g < s ? (float)mean[g] : 0.f;
#pragma unroll
    for (int r = ty; r < 64; r += 4) {
        const int64_t i = i0 + r;
        float v = 0.f;
        if (i < n && g < s) v = __ldg(P + i * ld + g) - mu;
        tile[r][tx] = v;
    }
    __syncthreads();
    // write transposed: row = gene (g0 + r), contiguous along spots (i0 + tx)
#pragma unroll
    for (int r = ty; r < 64; r += 4) {
        const float z = tile[tx][r];
        const __nv_bfloat16 hi = __float2bfloat16_rn(z);
        const __nv_bfloat16 lo = __float2bfloat16_rn(z - __bfloat162float(hi));
        const int64_t o = (g0 + r) * n_pad + i0 + tx;
        Zhi[o] = hi;
        Zlo[o] = lo;
    }
}

// the same for the super-tile kernel, in PANELS of 32 spots: panel b holds one 128-byte row per gene,
//   Zt[(b * s_pad + g) * 64 + 0..31] = hi,  [.. + 32..63] = lo   of spots 32 b .. 32 b + 31,
// so the 256 rows x 32 spots (hi and lo) a TMA box fetches are ONE contiguous 32 KB block of memory (one DRAM page, one TLB
// entry) instead of 256 lines 2.8 MB apart.  16-byte stores, 8 KB contiguous per (block, panel).
__global__ void __launch_bounds__(256)
center_split_interleave_kernel(const float* __restrict__ P, int64_t n, int64_t s, int64_t ld, const double* __restrict__ mean,
                               __nv_bfloat16* __restrict__ Zt, int64_t s_pad) {
    __shared__ float tile[64][65];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t g0 = (int64_t)blockIdx.x * 64, i0 = (int64_t)blockIdx.y * 64;
    const int64_t g = g0 + tx;
    const float mu = g < s ? (float)mean[g] : 0.f;
#pragma unroll
    for (int r = ty; r < 64; r += 4) {
        const int64_t i = i0 + r;
        float v = 0.f;
        if (i < n && g < s) v = __ldg(P + i * ld + g) - mu;
        tile[r][tx] = v;
    }
    __syncthreads();
    // a gene row of the block = 64 spots = one 128-byte line in each of two panels = 16 chunks of 8 bf16: chunk c = (panel, hi / lo, 8 spots)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int q = threadIdx.x + 256 * m;
        const int r = q >> 4, c = q & 15;
        const int kb = c >> 3, lo_part = (c >> 2) & 1, sub = c & 3;
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = tile[kb * 32 + sub * 8 + j][r];
            const __nv_bfloat16 hi = __float2bfloat16_rn(z);
            v[j] = lo_part ? __float2bfloat16_rn(z - __bfloat162float(hi)) : hi;
        }
        const int64_t o = (((i0 >> 5) + kb) * s_pad + g0 + r) * 64 + (c & 7) * 8;
        *reinterpret_cast<uint4*>(Zt + o) = *reinterpret_cast<const uint4*>(v);
    }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 Gram
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBN = 128, kBK = 64;           // tile (bf16: 64 elements = 128 B rows)
constexpr int kStages = 3;
constexpr int kTileBytes = kBM * kBK * 2;                // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;              // Hi, Hj, Li, Lj
constexpr int kGramThreads = 192;                        // warp0 TMA, warp1 MMA, warps 2-5 epilogue
constexpr int kTmemCols = 128;                           // fp32 accumulator 128 lanes x 128 columns
constexpr int kMaxChain = 8192;                          // spots per fp32 accumulation chain

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M=128 N=128 K=16, both operands K-major
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle, rows of 64 bf16: SBO = 8 rows x 128 B = 1024 B, LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B   (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t kIdescBf16F32_128x128 =
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__global__ void __launch_bounds__(kGramThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const int2* __restrict__ tiles, int n_kblocks_total, int kblocks_per_split,
                    float* __restrict__ partials) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int2 tile = tiles[blockIdx.x];                   // (i, j) block indices, i <= j
    const bool diag = tile.x == tile.y;
    const int kb0 = blockIdx.y * kblocks_per_split;
    const int kb1 = kb0 + kblocks_per_split < n_kblocks_total ? kb0 + kblocks_per_split : n_kblocks_total;
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {  // TMEM allocation is warp-collective
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], diag ? 2 * kTileBytes : 4 * kTileBytes);
                const int kc = (kb0 + kb) * kBK;
                tma_load_2d(st + 0 * kTileBytes, &map_hi, &full_bar[s], kc, tile.x * kBM);
                tma_load_2d(st + 2 * kTileBytes, &map_lo, &full_bar[s], kc, tile.x * kBM);
                if (!diag) {
                    tma_load_2d(st + 1 * kTileBytes, &map_hi, &full_bar[s], kc, tile.y * kBN);
                    tma_load_2d(st + 3 * kTileBytes, &map_lo, &full_bar[s], kc, tile.y * kBN);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t st = smem_u32(smem + s * kStageBytes);
                const uint32_t a_hi = st, b_hi = diag ? st : st + kTileBytes;
                const uint32_t a_lo = st + 2 * kTileBytes, b_lo = diag ? a_lo : st + 3 * kTileBytes;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint32_t ko = k * 32;  // 16 bf16 = 32 bytes along K inside the swizzle atom
                    const uint64_t dah = make_kmajor_sw128_desc(a_hi + ko), dbh = make_kmajor_sw128_desc(b_hi + ko);
                    const uint64_t dal = make_kmajor_sw128_desc(a_lo + ko), dbl = make_kmajor_sw128_desc(b_lo + ko);
                    umma_bf16(tmem_d, dah, dbh, kIdescBf16F32_128x128, (kb | k) ? 1u : 0u);
                    umma_bf16(tmem_d, dah, dbl, kIdescBf16F32_128x128, 1u);
                    umma_bf16(tmem_d, dal, dbh, kIdescBf16F32_128x128, 1u);
                }
                umma_commit(&empty_bar[s]);     // frees the smem stage when these MMAs retire
            }
            umma_commit(&accum_bar);            // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global partial tile =====
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                       // output row inside the tile
        float* out = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(kBM * kBN) + (size_t)row * kBN;
#pragma unroll 1
        for (int c0 = 0; c0 < kBN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float4* o4 = reinterpret_cast<float4*>(out + c0);
#pragma unroll
            for (int v = 0; v < 8; ++v)
                o4[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                    __uint_as_float(r[4 * v + 3]));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// 256 x 256 super-tiles: the default kernel.  gram_tcgen05_kernel above pulls 512 128-byte rows (H_i, L_i, H_j, L_j of 64 spots)
// through TMA for every 12 MMAs (768 tensor cycles) and is paced by those row requests (~3.2 cycles each per SM, the L2 -> SM
// cap: tensor pipe 48 % busy).  Here a CTA owns a 2 x 2 block of tiles and a stage holds 256 rows of each side over 32 spots,
// hi and lo interleaved in one 128-byte row (center_split_interleave_kernel): the same 512 row requests now feed four
// accumulators (4 x 128 TMEM columns = all 512) = 1536 tensor cycles.  Per 16-spot slice and accumulator row a:
//   D_a[128 x 256] += Hi_a Hj^T,  Hi_a Lj^T,  Li_a Hj^T      (N = 256 MMAs)
// - for each 128 x 128 tile the same slices in the same order as the tile kernel issues, so the partial tiles are
// bit-identical.  Diagonal super-tiles load one side and skip the sub-tile below the diagonal.
// ---------------------------------------------------------------------------------------------
constexpr int kSBM = 256, kSBK = 32;                    // super-tile edge; spots per stage
constexpr int kSOpBytes = kSBM * 128;                    // 32 KB: 256 rows x (32 hi + 32 lo) bf16
constexpr int kSStageBytes = 2 * kSOpBytes;              // side I, side J
constexpr int kSStages = 3;
constexpr int kSTmemCols = 512;
constexpr uint32_t kIdescBf16F32_128x256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(kGramThreads, 1)
gram_tcgen05_super_kernel(const __grid_constant__ CUtensorMap map_z, int nsb, int nb, int n_tiles, int n_kblocks_total /* panels of 32 spots */,
                          int kblocks_per_split, float* __restrict__ partials) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kSStages], empty_bar[kSStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // upper-triangular super-tile (I, J), I <= J, from the linear index
    int I = 0, rem = blockIdx.x;
    while (rem >= nsb - I) { rem -= nsb - I; ++I; }
    const int J = I + rem;
    const bool diag = I == J;
    const int kb0 = blockIdx.y * kblocks_per_split;
    const int kb1 = kb0 + kblocks_per_split < n_kblocks_total ? kb0 + kblocks_per_split : n_kblocks_total;
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kSStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kSTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kSStages;
                const uint32_t ph = (kb / kSStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * kSStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], diag ? kSOpBytes : 2 * kSOpBytes);
                const int row0 = (kb0 + kb) * (nb * kBM);                 // first row of the panel (s_pad rows each)
                tma_load_2d(st, &map_z, &full_bar[s], 0, row0 + I * kSBM);
                if (!diag) tma_load_2d(st + kSOpBytes, &map_z, &full_bar[s], 0, row0 + J * kSBM);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t kHalf = 128 * 128;           // byte offset of rows 128.. of a side
            constexpr uint32_t kLo = 64;                    // byte offset of the lo half inside a row
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kSStages;
                const uint32_t ph = (kb / kSStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a_s = smem_u32(smem + s * kSStageBytes);
                const uint32_t b_s = diag ? a_s : a_s + kSOpBytes;
#pragma unroll
                for (int k = 0; k < kSBK / 16; ++k) {
                    const uint32_t ko = k * 32;
                    const uint32_t acc = (kb | k) ? 1u : 0u;
                    const uint64_t dbh = make_kmajor_sw128_desc(b_s + ko), dbl = make_kmajor_sw128_desc(b_s + kLo + ko);
                    {   // accumulator row 0: all 256 columns
                        const uint64_t dah = make_kmajor_sw128_desc(a_s + ko), dal = make_kmajor_sw128_desc(a_s + kLo + ko);
                        umma_bf16(tmem_d, dah, dbh, kIdescBf16F32_128x256, acc);
                        umma_bf16(tmem_d, dah, dbl, kIdescBf16F32_128x256, 1u);
                        umma_bf16(tmem_d, dal, dbh, kIdescBf16F32_128x256, 1u);
                    }
                    const uint64_t dah = make_kmajor_sw128_desc(a_s + kHalf + ko), dal = make_kmajor_sw128_desc(a_s + kHalf + kLo + ko);
                    if (!diag) {
                        umma_bf16(tmem_d + 256, dah, dbh, kIdescBf16F32_128x256, acc);
                        umma_bf16(tmem_d + 256, dah, dbl, kIdescBf16F32_128x256, 1u);
                        umma_bf16(tmem_d + 256, dal, dbh, kIdescBf16F32_128x256, 1u);
                    } else {   // accumulator row 1 of a diagonal super-tile: only the sub-tile on the diagonal
                        const uint64_t dbh1 = make_kmajor_sw128_desc(b_s + kHalf + ko), dbl1 = make_kmajor_sw128_desc(b_s + kHalf + kLo + ko);
                        umma_bf16(tmem_d + 256 + 128, dah, dbh1, kIdescBf16F32_128x128, acc);
                        umma_bf16(tmem_d + 256 + 128, dah, dbl1, kIdescBf16F32_128x128, 1u);
                        umma_bf16(tmem_d + 256 + 128, dal, dbh1, kIdescBf16F32_128x128, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&accum_bar);
        }
    } else {
        mbar_wait(&accum_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;
        const int row = q * 32 + lane;
#pragma unroll 1
        for (int ab = 0; ab < 4; ++ab) {
            const int a = ab >> 1, b = ab & 1;
            const int i = 2 * I + a, j = 2 * J + b;
            if (i > j || j >= nb) continue;                  // below the diagonal / past the padded matrix
            const int t = i * nb - (i * (i - 1)) / 2 + (j - i);
            float* out = partials + ((size_t)blockIdx.y * n_tiles + t) * (size_t)(kBM * kBN) + (size_t)row * kBN;
#pragma unroll 1
            for (int c0 = 0; c0 < kBN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + b * 128 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4* o4 = reinterpret_cast<float4*>(out + c0);
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    o4[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                        __uint_as_float(r[4 * v + 3]));
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kSTmemCols) : "memory");
    }
}

// fixed-order fp64 sum of the split-K partial tiles, mirrored into the full symmetric matrix
__global__ void gram_reduce_kernel(const float* __restrict__ partials, const int2* __restrict__ tiles, int n_tiles, int ksplit,
                                   int64_t s, double* __restrict__ C) {
    const int t = blockIdx.x;
    const int2 tile = tiles[t];
    for (int e = threadIdx.x; e < kBM * kBN; e += blockDim.x) {
        const int r = e / kBN, c = e % kBN;
        const int64_t gr = (int64_t)tile.x * kBM + r, gc = (int64_t)tile.y * kBN + c;
        if (gr >= s || gc >= s) continue;
        if (tile.x == tile.y && c < r) continue;          // lower half of a diagonal tile comes from the mirror
        double a = 0.0;
        for (int k = 0; k < ksplit; ++k) a += (double)partials[((size_t)k * n_tiles + t) * (size_t)(kBM * kBN) + e];
        C[gr * s + gc] = a;
        C[gc * s + gr] = a;
    }
}

struct GramPlan {
    int64_t s_pad, n_pad;
    int nb, n_tiles, n_kblocks, kblocks_per_split, ksplit, cs_split;
    size_t off_zhi, off_zlo, off_tiles, off_part, off_cs, total;
};

static GramPlan gram_plan(int64_t n, int64_t s) {
    GramPlan p;
    p.s_pad = round_up<int64_t>(s, kBM);
    p.n_pad = round_up<int64_t>(n, kBK);
    p.nb = (int)(p.s_pad / kBM);
    p.n_tiles = p.nb * (p.nb + 1) / 2;
    p.n_kblocks = (int)(p.n_pad / kBK);
    // enough CTAs for ~4 waves, chains of at most kMaxChain spots, at least 8 k-blocks each
    int want = (int)ceil_div<int64_t>(4 * kNumSMs, p.n_tiles);
    int min_split = (int)ceil_div<int64_t>(p.n_pad, kMaxChain);
    int max_split = p.n_kblocks / 8 > 0 ? p.n_kblocks / 8 : 1;
    int ks = want > min_split ? want : min_split;
    if (ks > max_split) ks = max_split;
    if (ks < min_split) ks = min_split;
    if (ks < 1) ks = 1;
    p.kblocks_per_split = (int)ceil_div<int64_t>(p.n_kblocks, ks);
    p.ksplit = (int)ceil_div<int64_t>(p.n_kblocks, p.kblocks_per_split);
    p.cs_split = (int)ceil_div<int64_t>(n, kCsRows);
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 1024); return r; };
    p.off_zhi = take((size_t)p.s_pad * p.n_pad * 2);
    p.off_zlo = take((size_t)p.s_pad * p.n_pad * 2);
    p.off_tiles = take((size_t)p.n_tiles * sizeof(int2));
    p.off_part = take((size_t)p.ksplit * p.n_tiles * kBM * kBN * sizeof(float));
    p.off_cs = take((size_t)p.cs_split * s * sizeof(double));
    p.total = o;
    return p;
}

static int make_map(CUtensorMap* m, void* base, int64_t rows, int64_t cols_k, int box_k = kBK, int box_rows = kBM) {
    EncodeTiledFn enc = get_encode_tiled();
    CHR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {(cuuint64_t)cols_k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols_k * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CHR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

}  // namespace chr

extern "C" {

int chr_colsum_f64(const float* P, int64_t n, int64_t s, int64_t ld, double* sums, void* workspace, size_t workspace_bytes,
                   chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(P && sums && workspace && n > 0 && s > 0 && ld >= s, "chr_colsum_f64: bad arguments");
    const int nsplit = (int)ceil_div<int64_t>(n, kCsRows);
    CHR_REQUIRE(workspace_bytes >= (size_t)nsplit * s * sizeof(double) + 256, "chr_colsum_f64: workspace too small");
    double* part = (double*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    dim3 grid((unsigned)ceil_div<int64_t>(s, 128), (unsigned)nsplit);
    colsum_partial_kernel<<<grid, 128, 0, st>>>(P, n, s, ld, part);
    CHR_LAUNCH_CHECK();
    colsum_final_kernel<<<(unsigned)ceil_div<int64_t>(s, 128), 128, 0, st>>>(part, nsplit, s, sums);
    CHR_LAUNCH_CHECK();
    return 0;
}

size_t chr_colsum_workspace_bytes(int64_t n, int64_t s) {
    return (size_t)chr::ceil_div<int64_t>(n > 0 ? n : 1, chr::kCsRows) * (size_t)(s > 0 ? s : 1) * sizeof(double) + 512;
}

size_t chr_gram_workspace_bytes(int64_t n, int64_t s) {
    if (n <= 0 || s <= 0) return 1024;
    return chr::gram_plan(n, s).total + 2048;
}

int chr_gram_centered_bf16x3(const float* P, int64_t n, int64_t s, int64_t ld, const double* mean, double* C_out, unsigned flags,
                             void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(P && mean && C_out && workspace, "chr_gram_centered_bf16x3: null pointer argument");
    CHR_REQUIRE(n > 1 && s > 0 && ld >= s, "chr_gram_centered_bf16x3: bad shape (n=%lld s=%lld ld=%lld)", (long long)n, (long long)s,
                (long long)ld);
    GramPlan p = gram_plan(n, s);
    CHR_REQUIRE(workspace_bytes >= p.total + 1024, "chr_gram_centered_bf16x3: workspace too small (%zu < %zu)", workspace_bytes,
                p.total + 1024);
    char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* zhi = (__nv_bfloat16*)(ws + p.off_zhi);
    __nv_bfloat16* zlo = (__nv_bfloat16*)(ws + p.off_zlo);
    int2* tiles = (int2*)(ws + p.off_tiles);
    float* part = (float*)(ws + p.off_part);
    const bool timing = flags & CHR_FLAG_TIMING;

    // upper-triangular tile list (host -> device; a few hundred bytes, pageable copy is synchronous wrt host memory)
    {
        static thread_local int2 host_tiles[4096];
        CHR_REQUIRE(p.n_tiles <= 4096, "chr_gram_centered_bf16x3: too many tiles (s too large)");
        int t = 0;
        for (int i = 0; i < p.nb; ++i)
            for (int j = i; j < p.nb; ++j) host_tiles[t++] = make_int2(i, j);
        CHR_CUDA(cudaMemcpyAsync(tiles, host_tiles, (size_t)p.n_tiles * sizeof(int2), cudaMemcpyHostToDevice, st));
    }
    dim3 tgrid((unsigned)(p.s_pad / 64), (unsigned)(p.n_pad / 64));
    const bool super = ((flags >> CHR_VARIANT_SHIFT) & 0xffu) != 2u;     // variant 2: the 128 x 128 tile kernel
    CUtensorMap map_hi, map_lo;
    if (super) {
        // zhi and zlo are adjacent in the workspace: one array of n_pad / 32 panels x s_pad rows x 128 bytes
        CHR_REQUIRE(p.off_zlo == p.off_zhi + (size_t)p.s_pad * p.n_pad * 2, "chr_gram_centered_bf16x3: operand buffers are not adjacent");
        CHR_REQUIRE((p.n_pad / kSBK) * p.s_pad < (1LL << 31), "chr_gram_centered_bf16x3: matrix too large for the panel layout");
        center_split_interleave_kernel<<<tgrid, 256, 0, st>>>(P, n, s, ld, mean, zhi, p.s_pad);
        CHR_LAUNCH_CHECK();
        if (make_map(&map_hi, zhi, (p.n_pad / kSBK) * p.s_pad, 64, 64, kSBM)) return 3;
        const size_t smem = (size_t)kSStages * kSStageBytes + 1024;
        CHR_CUDA(cudaFuncSetAttribute(gram_tcgen05_super_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nsb = (p.nb + 1) / 2;
        dim3 grid((unsigned)(nsb * (nsb + 1) / 2), (unsigned)p.ksplit);
        timer_begin(CHR_FAMILY_GRAM, timing, st);
        gram_tcgen05_super_kernel<<<grid, kGramThreads, smem, st>>>(map_hi, nsb, p.nb, p.n_tiles, p.n_kblocks * (kBK / kSBK),
                                                                    p.kblocks_per_split * (kBK / kSBK), part);
        timer_end(CHR_FAMILY_GRAM, timing, st);
    } else {
        center_split_transpose_kernel<<<tgrid, 256, 0, st>>>(P, n, s, ld, mean, zhi, zlo, p.n_pad);
        CHR_LAUNCH_CHECK();
        if (make_map(&map_hi, zhi, p.s_pad, p.n_pad)) return 3;
        if (make_map(&map_lo, zlo, p.s_pad, p.n_pad)) return 3;
        const size_t smem = (size_t)kStages * kStageBytes + 1024;
        CHR_CUDA(cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)p.n_tiles, (unsigned)p.ksplit);
        timer_begin(CHR_FAMILY_GRAM, timing, st);
        gram_tcgen05_kernel<<<grid, kGramThreads, smem, st>>>(map_hi, map_lo, tiles, p.n_kblocks, p.kblocks_per_split, part);
        timer_end(CHR_FAMILY_GRAM, timing, st);
    }
    CHR_LAUNCH_CHECK();
    gram_reduce_kernel<<<(unsigned)p.n_tiles, 256, 0, st>>>(part, tiles, p.n_tiles, p.ksplit, s, C_out);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
