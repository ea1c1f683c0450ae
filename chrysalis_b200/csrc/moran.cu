// K3/K4 - Moran's I (and Geary's C) for every gene of a dense spots x genes matrix, one pass.
//
// Replaces the per-gene loop  esda.moran.Moran(gene_matrix[c], w, permutations=0).I
// [+ esda.geary.Geary(...).C]  of /root/reference/chrysalis/core.py:71-76.
//
// Layout: X is [n_spots][ld] float32 row-major (row = spot, spots in a spatially coherent
// order).  A CTA owns a 128-gene column slab and a range of spots; a warp owns one row slab at
// a time (128 genes = one float4 per lane = one coalesced 512-byte segment), every lane
// accumulates its four genes over the spots it visits, so there is no cross-thread reduction in
// the streaming loop.
//
// Default kernel (moran_tile_kernel): the slab is streamed HBM -> shared memory in tiles of 112 spots x 128 genes
// (56 KB, three stages) by two producer warps (cp.async, one coalesced 512-byte segment per warp instruction, arriving
// on per-stage mbarriers - no CTA barrier in the loop) together with the tile's slice of the quad plan of W
// (moran_plan.cu).  14 compute warps walk the quads of the tile, 2 per warp: own rows and in-tile neighbour rows are
// read from the staged tile, the ~22 % out-of-tile rows through L1/L2 (issued first so that their latency hides behind
// the in-tile rounds).  HBM sees X once (ncu: 1.09 x the algorithmic bytes).
//
// Numerics: values are shifted by a per-gene pilot mean p (mean of a fixed
// spot sample) before any product, so the one-pass identities
//     z'z  = S2 - n m'^2,      k z'Wz = SW - m' (k S1 + SC) + m'^2 n k       (m' = S1/n)
// with S1 = sum x', S2 = sum x'^2, SW = x''A x', SC = 1'A x' = sum_j indeg_j x'_j
// carry no cancellation; partial sums are fp32 over the 16 spots a warp visits in two tiles, then fp32
// (shared memory) over the CTA's spot range, then fp64 across warps and CTAs in fixed order.
#include <vector>

#include "async.cuh"
#include "moran.cuh"

namespace chr {

constexpr int kMoranThreads = 256;
constexpr int kMoranWarps = kMoranThreads / 32;
constexpr int kPilotSample = 512;
constexpr int kPilotSplit = 8;

// ---- pilot shift: mean of a fixed sample of spots, per gene (fixed summation order) -----------
__global__ void __launch_bounds__(64 * kPilotSplit)
moran_pilot_kernel(const float* __restrict__ X, int64_t n, int64_t ld, int64_t ncol, float* __restrict__ pilot) {
    __shared__ float red[kPilotSplit][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + tx;
    const int ns = n < kPilotSample ? (int)n : kPilotSample;
    float s = 0.f;
    if (col < ncol) {
#pragma unroll 8
        for (int t = ty; t < ns; t += kPilotSplit) {
            const int64_t i = (int64_t)t * n / ns;
            s += __ldg(X + i * ld + col);
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < kPilotSplit; ++y) a += red[y][tx];
        pilot[col] = a / (float)ns;
    }
}

// same per row block (block-diagonal input, blockIdx.y = block): a fixed sample of the block's REAL rows, so every sample
// is shifted by a value near its own mean and the padding rows never enter
__global__ void __launch_bounds__(64 * kPilotSplit)
moran_pilot_blocks_kernel(const float* __restrict__ X, int64_t ld, int64_t ncol, const int64_t* __restrict__ block_rows,
                          const int64_t* __restrict__ block_n, int64_t gpad, float* __restrict__ pilot_base) {
    __shared__ float red[kPilotSplit][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + tx;
    const int64_t r0 = block_rows[blockIdx.y], nb = block_n[blockIdx.y];
    const int ns = nb < kPilotSample ? (int)nb : kPilotSample;
    float* __restrict__ pilot = pilot_base + (int64_t)blockIdx.y * gpad;
    float s = 0.f;
    if (col < ncol) {
#pragma unroll 8
        for (int t = ty; t < ns; t += kPilotSplit) s += __ldg(X + (r0 + (int64_t)t * nb / ns) * ld + col);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < kPilotSplit; ++y) a += red[y][tx];
        pilot[col] = a / (float)ns;
    }
}

// =============================================================================================
// per-spot stream kernel (variant 4; fallback for k > 64).  W's neighbour lists are staged in
// shared memory 128 spots at a time by cp.async; every warp works on 2 spots at once (18
// independent 512-byte row loads in flight); packed f32x2 arithmetic; rows served by L1/L2.
// =============================================================================================
// one elected lane asks the TMA unit to pull a row segment into L2 ahead of use
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

constexpr int kV2Chunk = 128;                 // spots whose neighbour lists are staged at once
constexpr int kV2PerWarp = kV2Chunk / kMoranWarps;  // 16 = level-1 accumulation length
constexpr int kV2SpotsPerCta = 2048;          // 16 chunks -> level-2 accumulation length 16

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int K, bool GEARY, int MINB, bool SHIFT>
__global__ void __launch_bounds__(kMoranThreads, MINB)
moran_stream_kernel(const float* __restrict__ X, int64_t n, uint32_t row_bytes, int64_t ncol, const int32_t* __restrict__ nbr,
                    int k_rt, int nbr_vec16, const float* __restrict__ pilot, double* __restrict__ part, int64_t gpad) {
    const int k = K ? K : k_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* s_idx = reinterpret_cast<int32_t*>(smem_raw);           // [2][kV2Chunk * k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile_col = (int64_t)blockIdx.x * kGeneTile;
    int64_t col = tile_col + lane * 4;
    if (col >= ncol) col = tile_col;                                  // idle lanes shadow lane 0 (results unused)
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int64_t s_begin = (int64_t)blockIdx.y * kV2SpotsPerCta;
    const int64_t s_end = s_begin + kV2SpotsPerCta < n ? s_begin + kV2SpotsPerCta : n;
    const uint32_t tile_bytes = (uint32_t)((ncol - tile_col < kGeneTile ? ncol - tile_col : kGeneTile) * 4);

    Row4 p;
    p.a = p.b = f2_pack(0.f, 0.f);
    if (SHIFT) {
        const float4 pf = __ldg(reinterpret_cast<const float4*>(pilot + col));
        p.a = f2_pack(pf.x, pf.y);
        p.b = f2_pack(pf.z, pf.w);
    }

    const f2_t z2 = f2_pack(0.f, 0.f);
    Row4 l1[kNStat], l2[kNStat];
#pragma unroll
    for (int s = 0; s < kNStat; ++s) { l1[s].a = l1[s].b = z2; l2[s].a = l2[s].b = z2; }

    auto stage = [&](int buf, int64_t c0) {   // neighbour lists of spots [c0, c0 + chunk) -> smem
        int32_t* dst = s_idx + buf * kV2Chunk * k;
        const int64_t rem = (n - c0 < kV2Chunk ? n - c0 : kV2Chunk) * k;   // valid ints
        const int32_t* src = nbr + c0 * k;
        if (nbr_vec16) {
            for (int t = threadIdx.x * 4; t < rem; t += kMoranThreads * 4) cp_async_16(dst + t, src + t);   // rem % 4 == 0 when k % 4 == 0, else scalar path
        } else {
            for (int t = threadIdx.x; t < rem; t += kMoranThreads) cp_async_4(dst + t, src + t);
        }
        cp_async_commit();
    };

    auto accumulate = [&](Row4 x, Row4 s) {   // s is already a sum of shifted values
        if (SHIFT) r4_sub(x, p);
        r4_add(l1[0], x);
        r4_fma(l1[1], x, x);
        r4_fma(l1[2], x, s);
        r4_add(l1[3], s);
    };

    stage(0, s_begin);
    int buf = 0;
    for (int64_t c0 = s_begin; c0 < s_end; c0 += kV2Chunk, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();                                              // chunk `buf` visible; previous chunk fully consumed
        if (c0 + kV2Chunk < s_end) stage(buf ^ 1, c0 + kV2Chunk);
        const int32_t* my_idx = s_idx + buf * kV2Chunk * k + warp * kV2PerWarp * k;
        const int64_t w0 = c0 + warp * kV2PerWarp;
        // pull this warp's next chunk of own rows towards L2 while the current one is consumed
        if (lane < kV2PerWarp) {
            const int64_t ip = w0 + kV2Chunk + lane;
            if (ip < s_end) l2_prefetch_bulk(reinterpret_cast<const char*>(X + tile_col) + (uint64_t)ip * row_bytes, tile_bytes);
        }
#pragma unroll 1
        for (int it = 0; it < kV2PerWarp; it += 2) {
            const int64_t i0 = w0 + it, i1 = i0 + 1;
            if (i0 >= s_end) break;
            const bool has1 = i1 < s_end;
            const uint32_t r0 = (uint32_t)i0, r1 = (uint32_t)(has1 ? i1 : i0);
            const int32_t* idx0 = my_idx + it * k;
            const int32_t* idx1 = has1 ? idx0 + k : idx0;
            Row4 x0 = ld_row4(base + (uint64_t)r0 * row_bytes);
            Row4 x1 = ld_row4(base + (uint64_t)r1 * row_bytes);
            Row4 s0, s1;
            s0.a = s0.b = s1.a = s1.b = z2;
            if constexpr (K > 0) {
                Row4 v0[K], v1[K];
                int ia[K], ib[K];
                if constexpr (K % 4 == 0) {   // neighbour lists as 128-bit broadcast loads
#pragma unroll
                    for (int j = 0; j < K; j += 4) {
                        const int4 qa = *reinterpret_cast<const int4*>(idx0 + j), qb = *reinterpret_cast<const int4*>(idx1 + j);
                        ia[j] = qa.x; ia[j + 1] = qa.y; ia[j + 2] = qa.z; ia[j + 3] = qa.w;
                        ib[j] = qb.x; ib[j + 1] = qb.y; ib[j + 2] = qb.z; ib[j + 3] = qb.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) { ia[j] = idx0[j]; ib[j] = idx1[j]; }
                }
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    v0[j] = ld_row4(base + (uint64_t)(uint32_t)ia[j] * row_bytes);
                    v1[j] = ld_row4(base + (uint64_t)(uint32_t)ib[j] * row_bytes);
                }
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (GEARY) { r4_sqdiff(l1[4], x0, v0[j]); if (has1) r4_sqdiff(l1[4], x1, v1[j]); }
                    if (SHIFT) { r4_sub(v0[j], p); r4_sub(v1[j], p); }   // shift before summing: partial sums stay O(std)
                    r4_add(s0, v0[j]);
                    r4_add(s1, v1[j]);
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < k; ++j) {
                    Row4 v0 = ld_row4(base + (uint64_t)(uint32_t)idx0[j] * row_bytes);
                    Row4 v1 = ld_row4(base + (uint64_t)(uint32_t)idx1[j] * row_bytes);
                    if (GEARY) { r4_sqdiff(l1[4], x0, v0); if (has1) r4_sqdiff(l1[4], x1, v1); }
                    if (SHIFT) { r4_sub(v0, p); r4_sub(v1, p); }
                    r4_add(s0, v0);
                    r4_add(s1, v1);
                }
            }
            accumulate(x0, s0);
            if (has1) accumulate(x1, s1);
        }
        // level-1 -> level-2 (every 16 spots of this warp)
#pragma unroll
        for (int s = 0; s < kNStat; ++s) {
            if (s == 4 && !GEARY) break;
            r4_add(l2[s], l1[s]);
            l1[s].a = l1[s].b = z2;
        }
    }

    // level-2 -> fp64, cross-warp reduction in fixed order, one statistic at a time
    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_raw);               // [kMoranWarps][kGeneTile] (reuses the index buffers)
#pragma unroll
    for (int st = 0; st < kNStat; ++st) {
        if (st == 4 && !GEARY) break;
        float f0, f1, f2, f3;
        f2_unpack(l2[st].a, f0, f1);
        f2_unpack(l2[st].b, f2, f3);
        red[warp * kGeneTile + lane * 4 + 0] = (double)f0;
        red[warp * kGeneTile + lane * 4 + 1] = (double)f1;
        red[warp * kGeneTile + lane * 4 + 2] = (double)f2;
        red[warp * kGeneTile + lane * 4 + 3] = (double)f3;
        __syncthreads();
        if (threadIdx.x < kGeneTile) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kMoranWarps; ++w) t += red[w * kGeneTile + threadIdx.x];
            part[((int64_t)blockIdx.y * kNStat + st) * gpad + tile_col + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// =============================================================================================
// tile kernel (default): 128-spot x 128-gene tiles staged in shared memory + the quad plan of W
// =============================================================================================
// A CTA has 14 compute warps and 2 producer warps.  For every tile c of the CTA's spot range
//   * a producer warp waits until stage c % 3 has been released by all compute warps (tile c - 3
//     consumed), then issues its half of tile c right away: 56 rows by cp.async (one coalesced
//     512-byte segment per warp instruction, zero fill outside the matrix) and half of the tile's
//     plan slice; the copies arrive on the stage's "full" mbarrier when they land
//   * a compute warp waits for tile c's "full" barrier, walks its 2 quads out of shared memory and
//     arrives on tile c's "empty" barrier.
// Two tiles are always in flight per SM, the loads of a tile start the moment its stage is free
// (not when some compute warp gets round to it), and there is no CTA-wide barrier in the loop.
// (A TMA tensor-map producer was measured first: with 512-byte box rows it delivered a tile too
//  slowly to keep the compute warps fed - gpurun_out r1k/r1n; plain 512-byte LDG streams reach 7.3 TB/s.)
constexpr int kTileWarps = 14;                                   // compute warps
constexpr int kTileProducers = 2;                                // producer warps (warps 14, 15)
constexpr int kTileThreads = 32 * (kTileWarps + kTileProducers);
constexpr int kTileQuadsPerWarp = kTileQuads / kTileWarps;       // 2
constexpr int kTileRowsPerProducer = kTileSpots / kTileProducers;   // 56 rows of every tile are fetched by each producer warp
static_assert(kTileQuads == kTileWarps * kTileQuadsPerWarp, "a tile is 2 quads per compute warp");
constexpr int kTileBytes = kTileSpots * kGeneTile * 4;           // 64 KB
constexpr int kTileMaxStages = 3;                                // 3 stages: two tiles in flight while one is consumed
constexpr int kTileOffBytes = 256;                               // 36 ints of the record-offset table per tile

__host__ __device__ constexpr int tile_stage_bytes(int buf_units) { return kTileBytes + kTileOffBytes + ((buf_units * 16 + 127) & ~127); }

__device__ __forceinline__ Row4 lds_row4(uint32_t addr) {
    Row4 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.a), "=l"(r.b) : "r"(addr));
    return r;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
    int4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
// padded list slots have coefficient 0: their row load is skipped (predicated, no branch); the
// register keeps its previous finite content, which then contributes 0 * finite = 0
__device__ __forceinline__ void lds_row4_if(Row4& v, uint32_t addr, float c) {
    asm volatile("{\n"
        ".reg .pred q;\n"
        "setp.neu.f32 q, %3, 0f00000000;\n"
        "@q ld.shared.v2.b64 {%0, %1}, [%2];\n"
        "}\n"
        : "+l"(v.a), "+l"(v.b)
        : "r"(addr), "f"(c));
}
__device__ __forceinline__ void ldg_row4_if(Row4& v, const char* p, float c) {
    asm volatile("{\n"
        ".reg .pred q;\n"
        "setp.neu.f32 q, %3, 0f00000000;\n"
        "@q ld.global.nc.v2.b64 {%0, %1}, [%2];\n"
        "}\n"
        : "+l"(v.a), "+l"(v.b)
        : "l"(p), "f"(c));
}
// in-tile list slot: byte offset of the row inside the stage, negative for a padded slot (load skipped)
__device__ __forceinline__ void lds_row4_ifpos(Row4& v, uint32_t base, int off) {
    asm volatile("{\n"
        ".reg .pred q;\n"
        ".reg .b32 a;\n"
        "setp.ge.s32 q, %3, 0;\n"
        "add.u32 a, %2, %3;\n"
        "@q ld.shared.v2.b64 {%0, %1}, [a];\n"
        "}\n"
        : "+l"(v.a), "+l"(v.b)
        : "r"(base), "r"(off));
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a lost completion becomes a trap (an error code), never a hung GPU
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
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

// MAPPED: spot i reads row row_map[i] of X (permutation inference: z is permuted, W stays)
template <bool GEARY, bool SHIFT, bool MAPPED>
__global__ void __launch_bounds__(kTileThreads, 1)
moran_tile_kernel(const float* __restrict__ X, int n, uint32_t row_bytes, int ncol, const int32_t* __restrict__ row_map,
                  const int32_t* __restrict__ plan_off, const int4* __restrict__ plan_rec, int buf_units, int nstages, int tiles_per_cta,
                  const float* __restrict__ pilot_base, const int32_t* __restrict__ cta_block, double* __restrict__ part, int64_t gpad) {
    constexpr int NST = GEARY ? 5 : 4;
    // block-diagonal input: every row block (sample) has its own shift vector, the CTA takes its block's
    const float* __restrict__ pilot = pilot_base + (cta_block ? (int64_t)__ldg(cta_block + blockIdx.y) * gpad : 0);
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    uint32_t smem0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;            // everything below is addressed in the shared window
    asm volatile("" : "+r"(smem0));                                    // keep it in a register (do not recompute the window base in the loop)
    const uint32_t stage_bytes = (uint32_t)tile_stage_bytes(buf_units);
    const uint32_t full_bar = smem0 + nstages * stage_bytes, empty_bar = full_bar + 8 * kTileMaxStages;
    // level-2 accumulators live in shared memory (one float4 per statistic per compute thread): 16 registers less
    const uint32_t l2_lane = empty_bar + 8 * kTileMaxStages + threadIdx.x * 16;
    constexpr uint32_t kL2Stride = kTileWarps * 32 * 16;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.x * kGeneTile;
    const int nquads = (n + 3) >> 2;
    const int ntiles = (nquads + kTileQuads - 1) / kTileQuads;
    const int t_begin = blockIdx.y * tiles_per_cta;
    const int t_end = t_begin + tiles_per_cta < ntiles ? t_begin + tiles_per_cta : ntiles;
    const int my_tiles = t_end - t_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar + 8 * s), "r"(32 * kTileProducers) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_bar + 8 * s), "r"(kTileWarps) : "memory");
        }
        mbar_fence_init();
    }
    __syncthreads();

    int col = tile_col + lane * 4;
    const bool col_ok = col < ncol;
    if (!col_ok) col = tile_col;                                      // idle lanes shadow lane 0 (results unused)
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int col_bytes = col_ok ? 16 : 0;

    if (warp >= kTileWarps) {
        // ===== producer warps: stream the tiles and their plan slices, two tiles ahead of the compute warps =====
        const int pw = warp - kTileWarps;
        int row0 = t_begin * kTileSpots + pw * kTileRowsPerProducer;            // first spot this warp fetches of the tile
        const char* src = base + (uint64_t)(uint32_t)row0 * row_bytes;
        const uint64_t tile_stride = (uint64_t)kTileSpots * row_bytes;
        const uint32_t dst_lane = (uint32_t)(pw * kTileRowsPerProducer) * (kGeneTile * 4) + lane * 16;
        int avail = n - (t_begin * kTileSpots + pw * kTileRowsPerProducer);      // valid rows from this warp's first row on
        int u_start = __ldg(plan_off + t_begin * kTileQuads);
        int sidx = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int c = 0; c < my_tiles; ++c) {
            const int q1 = (t_begin + c + 1) * kTileQuads;
            const int u_end = __ldg(plan_off + (q1 < nquads ? q1 : nquads));
            if (c >= nstages) mbar_wait_u32(empty_bar + 8 * sidx, phase ^ 1u);     // tile c - nstages released by every compute warp
            const uint32_t st = smem0 + sidx * stage_bytes;
            if (MAPPED) {
                // source rows through the map: two coalesced loads cover the warp's 56 rows, then warp shuffles
                const int m0 = row0 + lane < n ? __ldg(row_map + row0 + lane) : 0;
                const int m1 = row0 + 32 + lane < n ? __ldg(row_map + row0 + 32 + lane) : 0;
#pragma unroll 8
                for (int r = 0; r < kTileRowsPerProducer; ++r) {
                    const int mr = __shfl_sync(0xffffffffu, r < 32 ? m0 : m1, r & 31);
                    cp_async_16_zfill(st + dst_lane + r * (kGeneTile * 4), base + (uint64_t)(uint32_t)mr * row_bytes, r < avail ? col_bytes : 0);
                }
            } else if (avail >= kTileRowsPerProducer) {
#pragma unroll 8
                for (int r = 0; r < kTileRowsPerProducer; ++r)
                    cp_async_16_zfill(st + dst_lane + r * (kGeneTile * 4), src + (uint64_t)r * row_bytes, col_bytes);
            } else {
#pragma unroll 1
                for (int r = 0; r < kTileRowsPerProducer; ++r)
                    cp_async_16_zfill(st + dst_lane + r * (kGeneTile * 4), r < avail ? src + (uint64_t)r * row_bytes : base, r < avail ? col_bytes : 0);
            }
            const int q0 = (t_begin + c) * kTileQuads;
            const int pt = pw * 32 + lane;                                        // 0 .. 63 over both producer warps
            if (pt < 9) cp_async_16_zfill(st + kTileBytes + pt * 16, plan_off + q0 + pt * 4, 16);
            for (int u = pt; u < u_end - u_start; u += 32 * kTileProducers)
                cp_async_16_zfill(st + kTileBytes + kTileOffBytes + u * 16, plan_rec + u_start + u, 16);
            cp_async_mbar_arrive(full_bar + 8 * sidx);
            u_start = u_end;
            src += tile_stride;
            row0 += kTileSpots;
            avail -= kTileSpots;
            if (++sidx == nstages) { sidx = 0; phase ^= 1u; }
        }
    }

    Row4 p;
    p.a = p.b = f2_pack(0.f, 0.f);
    if (SHIFT) {
        const float4 pf = __ldg(reinterpret_cast<const float4*>(pilot + col));
        p.a = f2_pack(pf.x, pf.y);
        p.b = f2_pack(pf.z, pf.w);
    }
    const f2_t z2 = f2_pack(0.f, 0.f);
    Row4 l1[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) {
        l1[s].a = l1[s].b = z2;
        if (warp < kTileWarps) asm volatile("st.shared.v2.b64 [%0], {%1, %1};" ::"r"(l2_lane + s * kL2Stride), "l"(z2) : "memory");
    }
    Row4 va[4], vg[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) va[t].a = va[t].b = vg[t].a = vg[t].b = z2;

    // stage / phase of the tile being consumed, kept incrementally: no division in the loop
    int sidx = 0;
    uint32_t phase = 0;
    const int my_compute_tiles = warp < kTileWarps ? my_tiles : 0;
#pragma unroll 1
    for (int c = 0; c < my_compute_tiles; ++c) {
        mbar_wait_u32(full_bar + 8 * sidx, phase);
        const uint32_t st = smem0 + sidx * stage_bytes;
        uint32_t tile_lane = st + lane * 16;
        const uint32_t s_off = st + kTileBytes, s_rec = st + kTileBytes + kTileOffBytes;
        const int q0 = (t_begin + c) * kTileQuads;
        const int nq = nquads - q0 < kTileQuads ? nquads - q0 : kTileQuads;
        const int off0 = lds_i32(s_off);
        // per-tile addresses stay in registers (otherwise they are re-derived from the kernel arguments in every quad)
        uint32_t s_recbase = s_rec - (uint32_t)off0 * 16;
        asm volatile("" : "+r"(tile_lane), "+r"(s_recbase));
        static_assert(kTileQuadsPerWarp == 2, "both record offsets of the warp are fetched up front");
        const int offq0 = lds_i32(s_off + 8 * warp), offq1 = lds_i32(s_off + 8 * warp + 4);
#pragma unroll 1
        for (int qi = 0; qi < kTileQuadsPerWarp; ++qi) {
            const int ql = warp * kTileQuadsPerWarp + qi;
            if (ql >= nq) break;
            const uint32_t rec = s_recbase + (uint32_t)(qi == 0 ? offq0 : offq1) * 16;
            const int4 h0 = lds_i4(rec);
            const float4 h1 = lds_f4(rec + 16);                       // c12 c13 c23 -
            const int RL = h0.x & 0xff, RG = (h0.x >> 8) & 0xff, nvalid = (h0.x >> 16) & 0xff;
            const bool irregular = (h0.x >> 30) & 1;                  // some indeg != k in this quad
            Row4 x[4], s[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) s[t].a = s[t].b = z2;
            // the first out-of-tile round is issued before the in-tile rounds so its latency hides behind them
            const uint32_t grec = rec + 16 * kPlanHdrUnits;            // out-of-tile rounds first, then the in-tile rounds
            const uint32_t lrec = grec + 16 * kPlanRoundUnits * RG;
            float cg[4] = {0.f, 0.f, 0.f, 0.f};
            if (RG > 0) {
                const int4 rows = lds_i4(grec);
                const float4 cf = lds_f4(grec + 16);
                int rr[4] = {rows.x, rows.y, rows.z, rows.w};
                if (MAPPED) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) rr[t] = __ldg(row_map + rr[t]);
                }
                cg[0] = cf.x; cg[1] = cf.y; cg[2] = cf.z; cg[3] = cf.w;
#pragma unroll
                for (int t = 0; t < 4; ++t) ldg_row4_if(vg[t], base + (uint64_t)(uint32_t)rr[t] * row_bytes, cg[t]);
            }
            // a second out-of-tile round goes out now as well, into the registers the quad's own rows take later
            if (RG > 1) {
                const int4 rows = lds_i4(grec + 32);
                const float4 cf = lds_f4(grec + 48);
                int rr[4] = {rows.x, rows.y, rows.z, rows.w};
                if (MAPPED) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) rr[t] = __ldg(row_map + rr[t]);
                }
                const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    x[t].a = x[t].b = z2;
                    ldg_row4_if(x[t], base + (uint64_t)(uint32_t)rr[t] * row_bytes, cc[t]);
                }
            }
            // in-tile rounds; the plan words of round r + 1 are fetched while round r is multiplied (the record is
            // followed by at least one more 32-byte unit inside the staged slice or its padding)
            int4 offs = lds_i4(lrec);
#pragma unroll 1
            for (int r = 0; r < RL; ++r) {
                const int oo[4] = {offs.x, offs.y, offs.z, offs.w};
                const float4 cf = lds_f4(lrec + 32 * r + 16);
#pragma unroll
                for (int t = 0; t < 4; ++t) lds_row4_ifpos(va[t], tile_lane, oo[t]);
                offs = lds_i4(lrec + 32 * r + 32);
                const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc[t], va[t]);
            }
            // own rows, pilot shift and the pairs inside the quad
            auto own_rows = [&]() {
#pragma unroll
                for (int t = 0; t < 4; ++t) x[t] = lds_row4(tile_lane + (uint32_t)(4 * ql + t) * (kGeneTile * 4));
                if (SHIFT) {
                    const float4 cs = lds_f4(rec + 32);               // csum
                    const float ncs[4] = {-cs.x, -cs.y, -cs.z, -cs.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        r4_sub(x[t], p);
                        r4_fmac(s[t], ncs[t], p);
                    }
                }
                r4_fmac(s[0], __int_as_float(h0.y), x[1]);
                r4_fmac(s[0], __int_as_float(h0.z), x[2]);
                r4_fmac(s[0], __int_as_float(h0.w), x[3]);
                r4_fmac(s[1], h1.x, x[2]);
                r4_fmac(s[1], h1.y, x[3]);
                r4_fmac(s[2], h1.z, x[3]);
            };
            if (RG > 0) {
#pragma unroll
                for (int t = 0; t < 4; ++t) r4_fmac(s[t], cg[t], vg[t]);
                if (RG > 1) {
                    const float4 cf = lds_f4(grec + 48);
                    const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc[t], x[t]);
                }
#pragma unroll 1
                for (int r = 2; r < RG; ++r) {
                    const int4 rows = lds_i4(grec + 32 * r);
                    const float4 cf = lds_f4(grec + 32 * r + 16);
                    int rr[4] = {rows.x, rows.y, rows.z, rows.w};
                    if (MAPPED) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) rr[t] = __ldg(row_map + rr[t]);
                    }
                    const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) ldg_row4_if(vg[t], base + (uint64_t)(uint32_t)rr[t] * row_bytes, cc[t]);
#pragma unroll
                    for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc[t], vg[t]);
                }
            }
            own_rows();
            const float4 dg = lds_f4(rec + 48);                       // indeg - k
            const float dgs[4] = {dg.x, dg.y, dg.z, dg.w};
            auto accumulate = [&](int t) {
                r4_add(l1[0], x[t]);
                r4_fma(l1[1], x[t], x[t]);
                r4_fma(l1[2], x[t], s[t]);
                if (irregular) {
                    // SC = sum_j indeg_j x_j = k S1 + sum_j (indeg_j - k) x_j: only the second part is accumulated, and
                    // only in quads that have a spot with indeg != k (lattice boundaries; everywhere for random kNN)
                    r4_fmac(l1[3], dgs[t], x[t]);
                    if (GEARY) {                                      // same I with or without Geary's C
                        Row4 dx;
                        const f2_t dd = f2_dup(dgs[t]);
                        dx.a = f2_fma(dd, x[t].a, z2);
                        dx.b = f2_fma(dd, x[t].b, z2);
                        r4_fma(l1[4], dx, x[t]);
                    }
                }
            };
            if ((h0.x >> 16) == 4) {                                  // 4 valid spots, every indeg == k: the common quad
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    r4_add(l1[0], x[t]);
                    r4_fma(l1[1], x[t], x[t]);
                    r4_fma(l1[2], x[t], s[t]);
                }
            } else if (nvalid == 4) {
#pragma unroll
                for (int t = 0; t < 4; ++t) accumulate(t);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t < nvalid) accumulate(t);
            }
        }
        // this warp is done with the stage
        __syncwarp();
        if (lane == 0) mbar_arrive_u32(empty_bar + 8 * sidx);
        if (++sidx == nstages) { sidx = 0; phase ^= 1; }
        // level-1 (16 spots: two tiles of this warp) -> level-2
        if (c & 1) {
#pragma unroll
            for (int stt = 0; stt < NST; ++stt) {
                Row4 acc = lds_row4(l2_lane + stt * kL2Stride);
                r4_add(acc, l1[stt]);
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(l2_lane + stt * kL2Stride), "l"(acc.a), "l"(acc.b) : "memory");
                l1[stt].a = l1[stt].b = z2;
            }
        }
    }
    Row4 l2[NST];
#pragma unroll
    for (int stt = 0; stt < NST; ++stt) {
        l2[stt] = l1[stt];
        if (warp < kTileWarps) r4_add(l2[stt], lds_row4(l2_lane + stt * kL2Stride));
    }

    // per-thread fp32 sums -> fp64, cross-warp reduction in fixed order, one statistic at a time
    __syncthreads();                                                  // every tile consumed: the stages are free
    double* red = reinterpret_cast<double*>(smem_dyn + (smem0 - smem_u32(smem_dyn)));   // [kTileWarps][kGeneTile]
#pragma unroll
    for (int stt = 0; stt < NST; ++stt) {
        if (warp < kTileWarps) {
            float f0, f1, f2, f3;
            f2_unpack(l2[stt].a, f0, f1);
            f2_unpack(l2[stt].b, f2, f3);
            red[warp * kGeneTile + lane * 4 + 0] = (double)f0;
            red[warp * kGeneTile + lane * 4 + 1] = (double)f1;
            red[warp * kGeneTile + lane * 4 + 2] = (double)f2;
            red[warp * kGeneTile + lane * 4 + 3] = (double)f3;
        }
        __syncthreads();
        if (threadIdx.x < kGeneTile) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kTileWarps; ++w) t += red[w * kGeneTile + threadIdx.x];
            part[((int64_t)blockIdx.y * kNStat + stt) * gpad + tile_col + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// ---- finalize: fixed-order sum of the split partials, then the statistic in fp64 ------------
constexpr int kFinSplit = 8;
__global__ void __launch_bounds__(64 * kFinSplit)
moran_finalize_kernel(const double* __restrict__ part, int nsplit, int64_t gpad, int64_t g, int64_t n, int k,
                      const float* __restrict__ pilot, bool geary, bool sym, double* __restrict__ out_I,
                      double* __restrict__ out_C, double* __restrict__ out_stats) {
    __shared__ double red[kFinSplit][kNStat][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + tx;
    const int ns = geary ? kNStat : kNStat - 1;
    double S[kNStat] = {0, 0, 0, 0, 0};
    if (col < g)
        for (int sp = ty; sp < nsplit; sp += kFinSplit)
            for (int st = 0; st < ns; ++st) S[st] += part[((int64_t)sp * kNStat + st) * gpad + col];
    for (int st = 0; st < kNStat; ++st) red[ty][st][tx] = S[st];
    __syncthreads();
    if (ty != 0 || col >= g) return;
    for (int st = 0; st < kNStat; ++st) {
        double a = 0.0;
        for (int y = 0; y < kFinSplit; ++y) a += red[y][st][tx];
        S[st] = a;
    }
    const double dn = (double)n, dk = (double)k;
    if (sym) {                                                    // the tile kernel accumulates the (indeg - k) parts only
        S[3] += dk * S[0];
        S[4] += dk * S[1];
    }
    const double m = S[0] / dn;                                   // mean of shifted values
    const double zz = S[1] - dn * m * m;                          // sum (x - mean)^2
    const double zwz = (S[2] - m * (dk * S[0] + S[3]) + m * m * dn * dk) / dk;  // z' W z, W = Adj / k
    // I = (n / s0) * z'Wz / z'z with s0 = n for row-standardised kNN weights
    out_I[col] = zwz / zz;
    // Geary numerator * k = sum_ij adj_ij (x_i - x_j)^2: accumulated directly (stream kernel) or from
    // k sum x^2 + sum_j indeg_j x_j^2 - 2 x'Ax (tile kernel; S[4] = sum_j indeg_j x_j^2)
    const double gnum = geary ? (sym ? dk * S[1] + S[4] - 2.0 * S[2] : S[4]) : 0.0;
    if (geary && out_C) out_C[col] = (dn - 1.0) * (gnum / dk) / (2.0 * dn * zz);
    if (out_stats) {
        out_stats[0 * g + col] = m + (double)pilot[col];   // pilot is all-zero for pre-shifted input
        out_stats[1 * g + col] = zz;
        out_stats[2 * g + col] = zwz;
        out_stats[3 * g + col] = gnum / dk;
    }
}

// block-diagonal W (one block per sample): the partials of the CTAs of block b, which cover whole blocks only, are summed
// per block; every block has its own n, mean and z'z - one Moran's I / Geary's C per (block, gene).
__global__ void __launch_bounds__(64 * kFinSplit)
moran_finalize_blocks_kernel(const double* __restrict__ part, const int32_t* __restrict__ block_split, const int64_t* __restrict__ block_n,
                             int64_t gpad, int64_t g, int k, bool geary, double* __restrict__ out_I, double* __restrict__ out_C) {
    __shared__ double red[kFinSplit][kNStat][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + tx;
    const int b = blockIdx.y;
    const int s0 = block_split[b], s1 = block_split[b + 1];
    const int ns = geary ? kNStat : kNStat - 1;
    double S[kNStat] = {0, 0, 0, 0, 0};
    if (col < g)
        for (int sp = s0 + ty; sp < s1; sp += kFinSplit)
            for (int st = 0; st < ns; ++st) S[st] += part[((int64_t)sp * kNStat + st) * gpad + col];
    for (int st = 0; st < kNStat; ++st) red[ty][st][tx] = S[st];
    __syncthreads();
    if (ty != 0 || col >= g) return;
    for (int st = 0; st < kNStat; ++st) {
        double a = 0.0;
        for (int y = 0; y < kFinSplit; ++y) a += red[y][st][tx];
        S[st] = a;
    }
    const double dn = (double)block_n[b], dk = (double)k;
    S[3] += dk * S[0];                                            // the tile kernel accumulates the (indeg - k) parts only
    S[4] += dk * S[1];
    const double m = S[0] / dn;
    const double zz = S[1] - dn * m * m;
    const double zwz = (S[2] - m * (dk * S[0] + S[3]) + m * m * dn * dk) / dk;
    out_I[(int64_t)b * g + col] = zwz / zz;
    if (geary && out_C) out_C[(int64_t)b * g + col] = (dn - 1.0) * ((dk * S[1] + S[4] - 2.0 * S[2]) / dk) / (2.0 * dn * zz);
}

// padding rows of a block take their block's pilot value: exactly neutral after the shift
__global__ void moran_fill_rows_kernel(float* __restrict__ X, int64_t ld, int64_t ncol4, const int32_t* __restrict__ rows, int64_t n_rows,
                                       const int64_t* __restrict__ block_rows, int n_blocks, int64_t gpad, const float* __restrict__ pilot_base) {
    const int64_t c4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= ncol4) return;
    for (int64_t a = blockIdx.y; a < n_rows; a += gridDim.y) {
        const int64_t r = rows[a];
        int b = 0;
        while (b + 1 < n_blocks && block_rows[b + 1] <= r) ++b;
        *reinterpret_cast<float4*>(X + r * ld + 4 * c4) = *reinterpret_cast<const float4*>(pilot_base + (int64_t)b * gpad + 4 * c4);
    }
}

// shared with moran_lattice.cu: `part` is [nsplit][kNStat][gpad] in the tile kernel's convention (x'Ax', (indeg - k) parts)
int launch_moran_finalize(const double* part, int nsplit, int64_t gpad, int64_t n_genes, int64_t n_spots, int k, const float* pilot,
                          bool geary, double* out_I, double* out_C, double* out_stats, cudaStream_t st) {
    moran_finalize_kernel<<<(unsigned)ceil_div<int64_t>(n_genes, 64), 64 * kFinSplit, 0, st>>>(part, nsplit, gpad, n_genes, n_spots, k, pilot,
                                                                                             geary, true, out_I, out_C, out_stats);
    CHR_LAUNCH_CHECK();
    return 0;
}

// spots per CTA of the tile kernel: depends on n only, so the summation tree of a gene is identical
// however genes are sharded.  Long CTAs amortise the pipeline fill; small n keeps enough CTAs.
static int tile_tiles_per_cta(int64_t n) {
    const int64_t ntiles = ceil_div<int64_t>(n, kTileSpots);
    if (ntiles >= 64 * 16) return 64;       // 8192 spots
    if (ntiles >= 16 * 8) return 16;        // 2048 spots
    return 4;
}

struct MoranWs {
    int64_t gpad;
    int nsplit, ntile;
    size_t pilot_off, part_off, total;
};

static MoranWs moran_ws(int64_t n, int64_t ncol, bool tile, int tiles_per_cta = 0) {
    MoranWs p;
    const int64_t spots_per_cta = tile ? (int64_t)(tiles_per_cta ? tiles_per_cta : tile_tiles_per_cta(n)) * kTileSpots : kV2SpotsPerCta;
    p.nsplit = (int)ceil_div<int64_t>(n, spots_per_cta);
    p.ntile = (int)ceil_div<int64_t>(ncol, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 256); return r; };
    p.pilot_off = take((size_t)p.gpad * sizeof(float));
    p.part_off = take((size_t)p.nsplit * kNStat * p.gpad * sizeof(double));
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_moran_workspace_bytes(int64_t n_spots, int64_t n_genes, int k, unsigned flags) {
    (void)flags; (void)k;
    if (n_spots <= 0 || n_genes <= 0) return 256;
    const int64_t ncol = chr::round_up<int64_t>(n_genes, 4);
    const size_t a = chr::moran_ws(n_spots, ncol, false).total, b = chr::moran_ws(n_spots, ncol, true).total;
    return (a > b ? a : b) + 512;
}

int chr_moran_dense_f32(const float* X, int64_t n_spots, int64_t n_genes, int64_t ld, const int32_t* nbr, int k,
                        const void* plan, int plan_chunk_units, const int32_t* row_map, unsigned flags, double* out_I,
                        double* out_C, double* out_stats, void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(n_spots > 1 && n_genes > 0, "chr_moran_dense_f32: need n_spots > 1 and n_genes > 0 (got %lld, %lld)",
                (long long)n_spots, (long long)n_genes);
    CHR_REQUIRE(k > 0 && k < n_spots, "chr_moran_dense_f32: need 0 < k < n_spots (k=%d)", k);
    CHR_REQUIRE(ld >= n_genes && ld % 4 == 0, "chr_moran_dense_f32: ld (%lld) must be >= n_genes and a multiple of 4", (long long)ld);
    CHR_REQUIRE(((uintptr_t)X & 15) == 0, "chr_moran_dense_f32: X must be 16-byte aligned");
    CHR_REQUIRE(X && out_I && workspace && (nbr || plan), "chr_moran_dense_f32: null pointer argument");
    CHR_REQUIRE(n_spots < (1LL << 31) && ld * 4 < (1LL << 32), "chr_moran_dense_f32: needs n_spots < 2^31, row bytes < 2^32");
    const bool geary = flags & CHR_FLAG_GEARY;
    CHR_REQUIRE(!geary || out_C, "chr_moran_dense_f32: CHR_FLAG_GEARY needs out_C");
    const bool timing = flags & CHR_FLAG_TIMING;

    // the kernels read round_up(n_genes, 4) <= ld columns
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    const unsigned variant = (flags >> CHR_VARIANT_SHIFT) & 0xffu;
    const bool preshifted = flags & CHR_FLAG_PRESHIFTED;
    // variant 0: tile kernel over the quad plan of W.  variant 4 / no plan: per-spot stream kernel over nbr
    bool tile = variant != 4 && plan != nullptr;
    // shared memory of the tile kernel: 3 (or 2) stages of {64 KB tile, record offsets, plan slice}; very long
    // neighbour lists do not fit and take the stream kernel
    int nstages = kTileMaxStages;
    auto tile_smem = [&](int ns) {
        return (size_t)1024 + (size_t)ns * tile_stage_bytes(plan_chunk_units) + 16 * kTileMaxStages + (size_t)(geary ? 5 : 4) * kTileWarps * 32 * 16;
    };
    if (tile && tile_smem(nstages) > 227 * 1024) nstages = 2;
    if (tile && tile_smem(nstages) > 227 * 1024 && nbr) tile = false;
    CHR_REQUIRE(tile || nbr, "chr_moran_dense_f32: the stream kernel needs the neighbour lists");
    CHR_REQUIRE(tile || !row_map, "chr_moran_dense_f32: row_map (permutation inference) needs the plan-based tile kernel");
    MoranWs p = moran_ws(n_spots, ncol, tile);
    CHR_REQUIRE(workspace_bytes >= p.total, "chr_moran_dense_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* pilot = (float*)(ws + p.pilot_off);
    double* part = (double*)(ws + p.part_off);

    if (preshifted) CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)p.gpad * sizeof(float), st));
    else moran_pilot_kernel<<<(unsigned)ceil_div<int64_t>(ncol, 64), 64 * kPilotSplit, 0, st>>>(X, n_spots, ld, ncol, pilot);
    CHR_LAUNCH_CHECK();

    dim3 grid((unsigned)p.ntile, (unsigned)p.nsplit);
    const uint32_t rb = (uint32_t)(ld * 4);
    if (tile) {
        CHR_REQUIRE(((uintptr_t)plan & 255) == 0 && plan_chunk_units > 0, "chr_moran_dense_f32: bad plan (build it with chr_moran_plan_build)");
        const int64_t nquads = (n_spots + 3) / 4;
        const char* pb = (const char*)plan;
        const int32_t* plan_off = (const int32_t*)(pb + kPlanHeaderBytes);
        const int4* plan_rec = (const int4*)(pb + kPlanHeaderBytes + round_up<size_t>((size_t)(nquads + 1 + kPlanOffPad) * 4, 256));
        const size_t smem = tile_smem(nstages);
        CHR_REQUIRE(smem <= 227 * 1024, "chr_moran_dense_f32: neighbour lists too long for the staged plan (%zu B of shared memory)", smem);
        const int tpc = tile_tiles_per_cta(n_spots);
        timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_TILE(GG, SH)                                                                                            \
    do {                                                                                                            \
        if (row_map) CHR_TILE_M(GG, SH, true); else CHR_TILE_M(GG, SH, false);                                     \
    } while (0)
#define CHR_TILE_M(GG, SH, MP)                                                                                      \
    do {                                                                                                            \
        auto kern = moran_tile_kernel<GG, SH, MP>;                                                                 \
        CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        kern<<<grid, kTileThreads, smem, st>>>(X, (int)n_spots, rb, (int)ncol, row_map, plan_off, plan_rec, plan_chunk_units, nstages, \
                                               tpc, pilot, nullptr, part, p.gpad);                                 \
    } while (0)
        if (geary && preshifted) CHR_TILE(true, false);
        else if (geary) CHR_TILE(true, true);
        else if (preshifted) CHR_TILE(false, false);
        else CHR_TILE(false, true);
#undef CHR_TILE
#undef CHR_TILE_M
    } else {
        const int vec16 = (k % 4 == 0) && (((uintptr_t)nbr & 15) == 0);
        size_t smem = (size_t)2 * kV2Chunk * k * sizeof(int32_t);
        if (smem < sizeof(double) * kMoranWarps * kGeneTile) smem = sizeof(double) * kMoranWarps * kGeneTile;
        CHR_REQUIRE(smem <= 200 * 1024, "chr_moran_dense_f32: k too large (%d)", k);
        timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_V2(KK, GG, MB, SH)                                                                                      \
    do {                                                                                                            \
        auto kern = moran_stream_kernel<KK, GG, MB, SH>;                                                           \
        if (smem > 48 * 1024) CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kMoranThreads, smem, st>>>(X, n_spots, rb, ncol, nbr, k, vec16, pilot, part, p.gpad);        \
    } while (0)
#define CHR_V2K(GG, MB, SH)                              \
    switch (k) {                                         \
        case 4: CHR_V2(4, GG, MB, SH); break;            \
        case 6: CHR_V2(6, GG, MB, SH); break;            \
        case 8: CHR_V2(8, GG, MB, SH); break;            \
        default: CHR_V2(0, GG, MB, SH); break;           \
    }
        if (geary && preshifted) { CHR_V2K(true, 1, false) }
        else if (geary) { CHR_V2K(true, 1, true) }
        else if (preshifted) { CHR_V2K(false, 2, false) }
        else { CHR_V2K(false, 2, true) }
#undef CHR_V2K
#undef CHR_V2
    }
    timer_end(CHR_FAMILY_MORAN, timing, st);
    CHR_LAUNCH_CHECK();

    moran_finalize_kernel<<<(unsigned)ceil_div<int64_t>(n_genes, 64), 64 * kFinSplit, 0, st>>>(
        part, p.nsplit, p.gpad, n_genes, n_spots, k, pilot, geary, tile, out_I, out_C, out_stats);
    CHR_LAUNCH_CHECK();
    return 0;
}


// scratch of the block-diagonal call behind the tile kernel's workspace: per-block shift vectors + three small tables
static size_t blockdiag_extra_bytes(int64_t gpad, int n_blocks, int nsplit) {
    return chr::round_up<size_t>((size_t)n_blocks * gpad * sizeof(float), 256) + chr::round_up<size_t>((size_t)(2 * n_blocks + 1) * 8, 256) +
           chr::round_up<size_t>((size_t)(n_blocks + 1 + nsplit) * 4, 256);
}

size_t chr_moran_blockdiag_workspace_bytes(int64_t n_rows, int64_t n_genes, int n_blocks) {
    if (n_rows <= 0 || n_genes <= 0 || n_blocks <= 0) return 256;
    const int64_t ncol = chr::round_up<int64_t>(n_genes, 4);
    chr::MoranWs p = chr::moran_ws(n_rows, ncol, true, CHR_BLOCK_ALIGN_ROWS / chr::kTileSpots);
    return chr::round_up<size_t>(p.total, 256) + blockdiag_extra_bytes(p.gpad, n_blocks, p.nsplit) + 1024;
}

int chr_moran_blockdiag_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int k, const void* plan, int plan_chunk_units,
                            const int64_t* block_rows_host, const int64_t* block_n_host, int n_blocks, const int32_t* pad_rows,
                            int64_t n_pad, unsigned flags, double* out_I, double* out_C, void* workspace, size_t workspace_bytes,
                            chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int kTpc = CHR_BLOCK_ALIGN_ROWS / kTileSpots;           // tiles per CTA: a CTA never straddles two blocks
    static_assert(kTpc * kTileSpots == CHR_BLOCK_ALIGN_ROWS, "block alignment is a whole number of tiles");
    CHR_REQUIRE(X && plan && block_rows_host && block_n_host && out_I && workspace, "chr_moran_blockdiag_f32: null pointer argument");
    CHR_REQUIRE(n_rows > 1 && n_genes > 0 && k > 0 && n_blocks > 0 && n_blocks < 65536, "chr_moran_blockdiag_f32: bad sizes");
    CHR_REQUIRE(ld >= n_genes && ld % 4 == 0 && ((uintptr_t)X & 15) == 0, "chr_moran_blockdiag_f32: X must be 16-byte aligned, ld %% 4 == 0, ld >= n_genes");
    CHR_REQUIRE(n_rows < (1LL << 31) && ld * 4 < (1LL << 32), "chr_moran_blockdiag_f32: needs n_rows < 2^31, row bytes < 2^32");
    CHR_REQUIRE(((uintptr_t)plan & 255) == 0 && plan_chunk_units > 0, "chr_moran_blockdiag_f32: bad plan (build it with chr_moran_plan_build)");
    CHR_REQUIRE(n_pad == 0 || pad_rows, "chr_moran_blockdiag_f32: pad_rows missing");
    CHR_REQUIRE(block_rows_host[0] == 0 && block_rows_host[n_blocks] == n_rows, "chr_moran_blockdiag_f32: block_rows must run from 0 to n_rows");
    const bool geary = flags & CHR_FLAG_GEARY, timing = flags & CHR_FLAG_TIMING;
    CHR_REQUIRE(!geary || out_C, "chr_moran_blockdiag_f32: CHR_FLAG_GEARY needs out_C");
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    MoranWs p = moran_ws(n_rows, ncol, true, kTpc);
    // host tables: [block_n | block_rows] (int64) and [first split of every block | block of every CTA row range] (int32)
    std::vector<int64_t> tab64((size_t)2 * n_blocks + 1);
    std::vector<int32_t> tab32((size_t)n_blocks + 1 + p.nsplit);
    for (int b = 0; b < n_blocks; ++b) {
        const int64_t r0 = block_rows_host[b], r1 = block_rows_host[b + 1];
        CHR_REQUIRE(r1 > r0 && r0 % CHR_BLOCK_ALIGN_ROWS == 0 && (r1 % CHR_BLOCK_ALIGN_ROWS == 0 || b == n_blocks - 1),
                    "chr_moran_blockdiag_f32: block %d must start and end on a multiple of %d rows", b, CHR_BLOCK_ALIGN_ROWS);
        CHR_REQUIRE(block_n_host[b] > 1 && block_n_host[b] <= r1 - r0, "chr_moran_blockdiag_f32: block %d holds %lld spots in %lld rows", b,
                    (long long)block_n_host[b], (long long)(r1 - r0));
        tab64[b] = block_n_host[b];
        tab64[n_blocks + b] = r0;
        tab32[b] = (int32_t)(r0 / CHR_BLOCK_ALIGN_ROWS);
        for (int64_t sp = r0 / CHR_BLOCK_ALIGN_ROWS; sp < ceil_div<int64_t>(r1, CHR_BLOCK_ALIGN_ROWS); ++sp) tab32[n_blocks + 1 + sp] = b;
    }
    tab64[2 * n_blocks] = n_rows;
    tab32[n_blocks] = p.nsplit;
    const size_t base = round_up<size_t>(p.total, 256);
    CHR_REQUIRE(workspace_bytes >= base + blockdiag_extra_bytes(p.gpad, n_blocks, p.nsplit) + 256, "chr_moran_blockdiag_f32: workspace too small");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* part = (double*)(ws + p.part_off);
    float* pilot = (float*)(ws + base);                               // [n_blocks][gpad]: one shift vector per block
    int64_t* tab64_dev = (int64_t*)((char*)pilot + round_up<size_t>((size_t)n_blocks * p.gpad * sizeof(float), 256));
    int32_t* tab32_dev = (int32_t*)((char*)tab64_dev + round_up<size_t>(tab64.size() * 8, 256));
    const int64_t* block_n_dev = tab64_dev;
    const int64_t* block_rows_dev = tab64_dev + n_blocks;
    const int32_t* split_dev = tab32_dev;
    const int32_t* cta_block_dev = tab32_dev + n_blocks + 1;
    CHR_CUDA(cudaMemcpyAsync(tab64_dev, tab64.data(), tab64.size() * 8, cudaMemcpyHostToDevice, st));
    CHR_CUDA(cudaMemcpyAsync(tab32_dev, tab32.data(), tab32.size() * 4, cudaMemcpyHostToDevice, st));
    CHR_CUDA(cudaStreamSynchronize(st));                              // the host tables live on this stack frame

    // every block is shifted by the mean of a fixed sample of ITS real rows (the statistics are per block, and samples of a
    // cohort can sit at different expression levels); the padding rows take their block's shift value
    CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)n_blocks * p.gpad * sizeof(float), st));
    moran_pilot_blocks_kernel<<<dim3((unsigned)ceil_div<int64_t>(ncol, 64), (unsigned)n_blocks), 64 * kPilotSplit, 0, st>>>(
        X, ld, ncol, block_rows_dev, block_n_dev, p.gpad, pilot);
    CHR_LAUNCH_CHECK();
    if (n_pad > 0) {
        dim3 gf((unsigned)ceil_div<int64_t>(ncol / 4, 128), (unsigned)(n_pad < 1024 ? n_pad : 1024));
        moran_fill_rows_kernel<<<gf, 128, 0, st>>>(X, ld, ncol / 4, pad_rows, n_pad, block_rows_dev, n_blocks, p.gpad, pilot);
        CHR_LAUNCH_CHECK();
    }
    int nstages = kTileMaxStages;
    auto tile_smem = [&](int ns) {
        return (size_t)1024 + (size_t)ns * tile_stage_bytes(plan_chunk_units) + 16 * kTileMaxStages + (size_t)(geary ? 5 : 4) * kTileWarps * 32 * 16;
    };
    if (tile_smem(nstages) > 227 * 1024) nstages = 2;
    const size_t smem = tile_smem(nstages);
    CHR_REQUIRE(smem <= 227 * 1024, "chr_moran_blockdiag_f32: neighbour lists too long for the staged plan (%zu B of shared memory)", smem);
    const int64_t nquads = (n_rows + 3) / 4;
    const char* pb = (const char*)plan;
    const int32_t* plan_off = (const int32_t*)(pb + kPlanHeaderBytes);
    const int4* plan_rec = (const int4*)(pb + kPlanHeaderBytes + round_up<size_t>((size_t)(nquads + 1 + kPlanOffPad) * 4, 256));
    dim3 grid((unsigned)p.ntile, (unsigned)p.nsplit);
    const uint32_t rb = (uint32_t)(ld * 4);
    timer_begin(CHR_FAMILY_MORAN, timing, st);
    if (geary) {
        auto kern = moran_tile_kernel<true, true, false>;
        CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kTileThreads, smem, st>>>(X, (int)n_rows, rb, (int)ncol, nullptr, plan_off, plan_rec, plan_chunk_units, nstages, kTpc, pilot,
                                               cta_block_dev, part, p.gpad);
    } else {
        auto kern = moran_tile_kernel<false, true, false>;
        CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kTileThreads, smem, st>>>(X, (int)n_rows, rb, (int)ncol, nullptr, plan_off, plan_rec, plan_chunk_units, nstages, kTpc, pilot,
                                               cta_block_dev, part, p.gpad);
    }
    timer_end(CHR_FAMILY_MORAN, timing, st);
    CHR_LAUNCH_CHECK();
    moran_finalize_blocks_kernel<<<dim3((unsigned)ceil_div<int64_t>(n_genes, 64), (unsigned)n_blocks), 64 * kFinSplit, 0, st>>>(
        part, split_dev, block_n_dev, p.gpad, n_genes, k, geary, out_I, out_C);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
