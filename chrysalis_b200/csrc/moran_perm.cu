// K3p - permutation inference of Moran's I:  esda.moran.Moran(y, w, permutations=P)  (the reference's library draws
// np.random.permutation(z) P times per gene and recomputes I; used at
// /root/reference/article/A2_human_lymph_node/morans_i.ipynb:186, BASELINE cfg3: 120k bins x 25k genes, P = 999).
//
// A draw permutes the SPOTS against a fixed W and is shared by all genes.  Permutations are never materialised: pi_draw(i)
// is a keyed Feistel function (perm.cuh) evaluated where a row is fetched.
//
//   * chr_perm_row_map: pi_draw as an int32 row map for chr_moran_dense_f32 (general W: one pass of the quad-plan kernel
//     per draw).
//   * chr_moran_perm_lattice_f32 (square lattice, k = 8): the stencil kernel of moran_lattice.cu with a producer that
//     fetches row pi(spot) instead of the spot's own row, launched over BATCHES of draws with the gene slab as the slowest
//     grid dimension: all (chunk, draw) CTAs of one 128-gene slab run back to back, so at cfg3 the slab (120k x 512 B =
//     61 MB) is read from HBM once per batch and served from L2 for the other draws.  z'z and the mean do not change under a
//     permutation, so a draw only accumulates z_pi' A z_pi (stencil part + the correction list); values are shifted by
//     the gene's exact mean from the observed pass.  Per gene the draws are folded on the device into
//     #{I_sim >= I}, sum I_sim, sum I_sim^2 - everything p_sim, EI_sim, VI_sim, z_sim need.
#include "moran_lattice.cuh"
#include "perm.cuh"

namespace chr {

__global__ void perm_row_map_kernel(int32_t* __restrict__ out, uint32_t n, uint64_t seed, uint32_t draw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PermKeys pk = perm_keys(seed, draw, n);
    out[i] = (int32_t)perm_eval(pk, i);
}

// composed row maps of a batch of draws: maps[d][r] = matrix row whose values the node stored at row r takes under draw
// draw0 + d (row of spot pi(spot at r)), -1 for filler rows.  One independent load per row is all the producers then need.
__global__ void perm_maps_kernel(const int32_t* __restrict__ spot_of_row, const int32_t* __restrict__ row_of_spot, int64_t n_rows,
                                 uint32_t n_spots, uint64_t seed, uint32_t draw0, int32_t* __restrict__ maps) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int s = spot_of_row[r];
    int out = -1;
    if (s >= 0) {
        const PermKeys pk = perm_keys(seed, draw0 + blockIdx.y, n_spots);
        out = row_of_spot[perm_eval(pk, (uint32_t)s)];
    }
    maps[(int64_t)blockIdx.y * n_rows + r] = out;
}

// shift = the gene's mean (float32), inv = 1 / (k z'z); accumulators zeroed
__global__ void perm_prepare_kernel(const double* __restrict__ stats, int64_t g, int64_t gpad, int k, float* __restrict__ shift,
                                    double* __restrict__ inv_kzz, long long* __restrict__ larger, double* __restrict__ s1, double* __restrict__ s2) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gpad) return;
    if (c < g) {
        shift[c] = (float)stats[c];
        inv_kzz[c] = 1.0 / ((double)k * stats[g + c]);
        larger[c] = 0;
        s1[c] = 0.0;
        s2[c] = 0.0;
    } else {
        shift[c] = 0.f;
    }
}

template <bool MASKED>
__device__ __forceinline__ void perm_step(uint32_t addr, bool accumulate, const Row4& p, const float (&m)[kLatR], Row4 (&cur)[kLatR + 1],
                                          Row4 (&nxt)[kLatR + 1], Row4 (&pc)[kLatR + 1], Row4& acc) {
#pragma unroll
    for (int r = 0; r <= kLatR; ++r) {
        nxt[r] = lat_lds_row4(addr + r * kLatRowBytes);
        r4_sub(nxt[r], p);
    }
    if (accumulate) {
#pragma unroll
        for (int r = 0; r < kLatR; ++r) {
            Row4 s = pc[r + 1];
            r4_add(s, nxt[r + 1]);
            r4_add(s, nxt[r]);
            if (MASKED) {
                const f2_t mm = f2_dup(m[r]);
                const f2_t z2 = f2_pack(0.f, 0.f);
                Row4 mc;
                mc.a = f2_fma(mm, cur[r].a, z2);
                mc.b = f2_fma(mm, cur[r].b, z2);
                r4_fma(acc, mc, s);
            } else {
                r4_fma(acc, cur[r], s);
            }
        }
    }
#pragma unroll
    for (int r = 1; r <= kLatR; ++r) {
        pc[r] = cur[r];
        r4_add(pc[r], nxt[r]);
    }
}

// blockIdx.x = chunk + n_chunks * (draw inside the batch), blockIdx.y = gene slab (slowest: L2 reuse across the draws)
// BULK: every gathered row travels as ONE 512-byte bulk copy (cp.async.bulk, the TMA engine's 1-D form: a lane issues the row it
// resolved, completion is counted in bytes on the stage's mbarrier) instead of 32 lanes x 16-byte cp.async.
template <bool BULK>
__global__ void __launch_bounds__(kLatThreads, 1)
moran_lattice_perm_kernel(const float* __restrict__ X, uint32_t row_bytes, int ncol, const int4* __restrict__ bands, int n_bands,
                          const int4* __restrict__ chunks, int n_chunks, const int32_t* __restrict__ maps, int64_t n_rows,
                          const float* __restrict__ shift, double* __restrict__ part, int nsplit, int64_t gpad) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    uint32_t smem0 = (smem_u32(smem_dyn) + 127u) & ~127u;
    asm volatile("" : "+r"(smem0));
    const uint32_t full_bar = smem0 + kLatStages * kLatStageBytes, empty_bar = full_bar + 8 * kLatStages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.y * kGeneTile;
    const int chunk_id = blockIdx.x % n_chunks, dloc = blockIdx.x / n_chunks;
    const int4 ch = __ldg(chunks + chunk_id);
    const int4 bd = __ldg(bands + ch.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kLatStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar + 8 * s), "r"(BULK ? kLatProducers : 32 * kLatProducers) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_bar + 8 * s), "r"(kLatWarps) : "memory");
        }
        mbar_fence_init();
    }
    __syncthreads();

    int col = tile_col + lane * 4;
    const bool col_ok = col < ncol;
    if (!col_ok) col = tile_col;
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int col_bytes = col_ok ? 16 : 0;
    const int n_steps = ch.z + 2;
    const int n_iters = (n_steps + kLatStageCols - 1) / kLatStageCols;

    if (warp >= kLatWarps) {
        // ===== producer warps: row t of a column comes from the row of spot pi(spot at that node) =====
        const int pw = warp - kLatWarps;
        const int32_t* __restrict__ map = maps + (int64_t)dloc * n_rows;
        const bool has_next = ch.x + 1 < n_bands;
        int4 nb = make_int4(0, 0, 0, 0);
        if (has_next) nb = __ldg(bands + ch.x + 1);
        const char* pil = reinterpret_cast<const char*>(shift + col);
        // BULK: whole-tile sources (the last slab of a matrix whose row stride is not a multiple of 128 floats is narrower)
        const char* __restrict__ tile_base = reinterpret_cast<const char*>(X + tile_col);
        const char* tile_pil = reinterpret_cast<const char*>(shift + tile_col);
        const uint32_t tile_bytes = row_bytes - (uint32_t)tile_col * 4u < (uint32_t)kLatRowBytes ? row_bytes - (uint32_t)tile_col * 4u : (uint32_t)kLatRowBytes;
        constexpr int kRowsPerProducer = (kLatSegRows + kLatProducers - 1) / kLatProducers;
        static_assert(kRowsPerProducer <= 32, "one lane resolves one row of the column");
        // source row of the node this lane resolves (row pw + 4 * lane of column j of the chunk), -1 = neutral (shift vector)
        auto resolve = [&](int j) -> int {
            const int xc = ch.y + j;
            const bool valid = xc >= 0 && xc < bd.z;
            const int t = pw + kLatProducers * lane;
            int stored = -1;                                             // matrix row that sits at this node
            if (t < bd.w) {
                if (valid) stored = bd.x + xc * bd.w + t;
            } else if (t == bd.w && has_next) {
                const int xn = bd.y + xc - nb.y;
                if (xn >= 0 && xn < nb.z) stored = nb.x + xn * nb.w;
            }
            return (stored >= 0 && j <= ch.z) ? __ldg(map + stored) : -1;
        };
        int sidx = 0;
        uint32_t phase = 0;
        int src_cur[kLatStageCols], src_nxt[kLatStageCols];
#pragma unroll
        for (int cc = 0; cc < kLatStageCols; ++cc) src_cur[cc] = resolve(cc - 1);
#pragma unroll 1
        for (int it = 0; it < n_iters; ++it) {
            // the map loads behind the NEXT stage's rows are in flight while this stage's copies are issued
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) src_nxt[cc] = resolve((it + 1) * kLatStageCols + cc - 1);
            if (it >= kLatStages) lat_mbar_wait(empty_bar + 8 * sidx, phase ^ 1u);
            if constexpr (BULK) {
                // lane l copies row t = pw + kLatProducers * l of every column of the stage: the row it resolved itself
                const int t = pw + kLatProducers * lane;
                const int my_rows = (kLatSegRows - pw + kLatProducers - 1) / kLatProducers;
                const int j_first = it * kLatStageCols - 1;
                const int ncols = ch.z - j_first + 1 < kLatStageCols ? ch.z - j_first + 1 : kLatStageCols;
                const uint32_t bar = full_bar + 8 * sidx;
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(my_rows * ncols) * tile_bytes) : "memory");
                __syncwarp();
                if (t < kLatSegRows) {
                    const uint32_t dst = smem0 + sidx * kLatStageBytes + t * kLatRowBytes;
#pragma unroll
                    for (int cc = 0; cc < kLatStageCols; ++cc) {
                        if (cc >= ncols) break;
                        const int srow = src_cur[cc];
                        const char* q = srow >= 0 ? tile_base + (uint64_t)(uint32_t)srow * row_bytes : tile_pil;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + cc * kLatSegBytes),
                                     "l"(q), "r"(tile_bytes), "r"(bar)
                                     : "memory");
                    }
                }
                if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
#pragma unroll
                for (int cc = 0; cc < kLatStageCols; ++cc) src_cur[cc] = src_nxt[cc];
                continue;
            }
            const uint32_t st = smem0 + sidx * kLatStageBytes + lane * 16;
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) {
                const int j = it * kLatStageCols + cc - 1;
                if (j > ch.z) break;
                const uint32_t dst = st + cc * kLatSegBytes;
#pragma unroll 4
                for (int l = 0; l < kRowsPerProducer; ++l) {
                    const int t = pw + kLatProducers * l;
                    const int srow = __shfl_sync(0xffffffffu, src_cur[cc], l);
                    if (t < kLatSegRows) {
                        const char* q = srow >= 0 ? base + (uint64_t)(uint32_t)srow * row_bytes : pil;
                        lat_cp_async_16(dst + t * kLatRowBytes, q, col_bytes);
                    }
                }
            }
            lat_cp_async_mbar_arrive(full_bar + 8 * sidx);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) src_cur[cc] = src_nxt[cc];
        }
    }

    const float4 pf = __ldg(reinterpret_cast<const float4*>(shift + col));
    Row4 p;
    p.a = f2_pack(pf.x, pf.y);
    p.b = f2_pack(pf.z, pf.w);
    const f2_t z2 = f2_pack(0.f, 0.f);
    Row4 l1, l2;
    l1.a = l1.b = l2.a = l2.b = z2;

    if (warp < kLatWarps) {
        Row4 ca[kLatR + 1], cb[kLatR + 1], pc[kLatR + 1];
#pragma unroll
        for (int r = 0; r <= kLatR; ++r) ca[r].a = ca[r].b = cb[r].a = cb[r].b = pc[r].a = pc[r].b = z2;
        float m[kLatR];
#pragma unroll
        for (int r = 0; r < kLatR; ++r) m[r] = warp * kLatR + r < bd.w ? 1.f : 0.f;
        const bool masked = bd.w < kLatBH;
        const uint32_t warp_off = (uint32_t)(warp * kLatR) * kLatRowBytes + lane * 16;
        int sidx = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_iters; ++it) {
            lat_mbar_wait(full_bar + 8 * sidx, phase);
            const uint32_t a0 = smem0 + sidx * kLatStageBytes + warp_off;
            const int j0 = it * kLatStageCols - 1;
            if (masked) {
                perm_step<true>(a0, j0 != 0, p, m, ca, cb, pc, l1);
                if (j0 + 1 <= ch.z) perm_step<true>(a0 + kLatSegBytes, j0 + 1 != 0, p, m, cb, ca, pc, l1);
            } else {
                perm_step<false>(a0, j0 != 0, p, m, ca, cb, pc, l1);
                if (j0 + 1 <= ch.z) perm_step<false>(a0 + kLatSegBytes, j0 + 1 != 0, p, m, cb, ca, pc, l1);
            }
            __syncwarp();
            if (lane == 0) lat_mbar_arrive(empty_bar + 8 * sidx);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
            if ((it & (kLatFlushStages - 1)) == kLatFlushStages - 1) {
                r4_add(l2, l1);
                l1.a = l1.b = z2;
            }
        }
        r4_add(l2, l1);
    }

    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_dyn + (smem0 - smem_u32(smem_dyn)));   // [kLatWarps][kGeneTile]
    if (warp < kLatWarps) {
        float f0, f1, f2, f3;
        f2_unpack(l2.a, f0, f1);
        f2_unpack(l2.b, f2, f3);
        red[warp * kGeneTile + lane * 4 + 0] = (double)f0;
        red[warp * kGeneTile + lane * 4 + 1] = (double)f1;
        red[warp * kGeneTile + lane * 4 + 2] = (double)f2;
        red[warp * kGeneTile + lane * 4 + 3] = (double)f3;
    }
    __syncthreads();
    if (threadIdx.x < kGeneTile) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLatWarps; ++w) t += red[w * kGeneTile + threadIdx.x];
        part[((int64_t)dloc * nsplit + chunk_id) * gpad + tile_col + threadIdx.x] = 2.0 * t;     // every unordered pair once
    }
}

// corrections D = A - B under the permutation: x'_pi(i) x'_pi(j) for the kind-0 entries (grid: slab, split, draw)
__global__ void __launch_bounds__(32 * kCorrWarps)
moran_lattice_perm_corr_kernel(const float* __restrict__ X, uint32_t row_bytes, int ncol, const int4* __restrict__ corr, int64_t n_corr,
                               const int32_t* __restrict__ maps, int64_t n_rows, const float* __restrict__ shift, double* __restrict__ part,
                               int nsplit, int split0, int64_t gpad) {
    __shared__ double red[kCorrWarps][kGeneTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.x * kGeneTile;
    int col = tile_col + lane * 4;
    if (col >= ncol) col = tile_col;
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const float4 pf = __ldg(reinterpret_cast<const float4*>(shift + col));
    const int32_t* __restrict__ map = maps + (int64_t)blockIdx.z * n_rows;
    const int64_t e0 = (int64_t)blockIdx.y * kCorrPerCta, e1 = e0 + kCorrPerCta < n_corr ? e0 + kCorrPerCta : n_corr;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t e = e0 + warp; e < e1; e += kCorrWarps) {
        const int4 en = __ldg(corr + e);
        if (en.w != 0) continue;
        const int ri = __ldg(map + en.x), rj = __ldg(map + en.y);
        const float c = __int_as_float(en.z);
        const float4 xi = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)ri * row_bytes));
        const float4 xj = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)rj * row_bytes));
        acc[0] += (double)c * (double)(xi.x - pf.x) * (double)(xj.x - pf.x);
        acc[1] += (double)c * (double)(xi.y - pf.y) * (double)(xj.y - pf.y);
        acc[2] += (double)c * (double)(xi.z - pf.z) * (double)(xj.z - pf.z);
        acc[3] += (double)c * (double)(xi.w - pf.w) * (double)(xj.w - pf.w);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) red[warp][lane * 4 + q] = acc[q];
    __syncthreads();
    if (threadIdx.x < kGeneTile) {
        double t = 0.0;
        for (int w = 0; w < kCorrWarps; ++w) t += red[w][threadIdx.x];
        part[((int64_t)blockIdx.z * nsplit + split0 + blockIdx.y) * gpad + tile_col + threadIdx.x] = t;
    }
}

// folds a batch of draws into the per-gene accumulators (one thread per gene, draws in order)
__global__ void perm_fold_kernel(const double* __restrict__ part, int n_draws, int nsplit, int64_t gpad, int64_t g,
                                 const double* __restrict__ inv_kzz, const double* __restrict__ I_obs, long long* __restrict__ larger,
                                 double* __restrict__ s1, double* __restrict__ s2) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g) return;
    const double inv = inv_kzz[c], obs = I_obs[c];
    long long cnt = larger[c];
    double a1 = s1[c], a2 = s2[c];
    for (int d = 0; d < n_draws; ++d) {
        double sw = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) sw += part[((int64_t)d * nsplit + sp) * gpad + c];
        const double I = sw * inv;
        cnt += I >= obs;
        a1 += I;
        a2 += I * I;
    }
    larger[c] = cnt;
    s1[c] = a1;
    s2[c] = a2;
}

struct PermWs {
    int64_t gpad;
    int ntile, ncorr_split, nsplit, batch;
    size_t lat_off, stats_off, shift_off, inv_off, part_off, maps_off, total;
};

static PermWs perm_ws(int64_t n_rows, int64_t n_genes, int n_chunks, int64_t n_corr, int permutations) {
    PermWs p;
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    p.ntile = (int)ceil_div<int64_t>(ncol, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    p.ncorr_split = (int)ceil_div<int64_t>(n_corr, kCorrPerCta);
    p.nsplit = n_chunks + p.ncorr_split;
    // draws per launch: enough (chunk, draw) CTAs per gene slab to fill the GPU a few times, partials below ~256 MB
    int b = (int)ceil_div<int64_t>(4 * kNumSMs, n_chunks);
    if (b < 8) b = 8;
    if (b > 64) b = 64;
    while (b > 1 && (size_t)b * p.nsplit * p.gpad * 8 > ((size_t)256 << 20)) b /= 2;
    if (b > permutations) b = permutations > 0 ? permutations : 1;
    p.batch = b;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += round_up<size_t>(bytes, 256); return r; };
    p.lat_off = take(chr_moran_lattice_workspace_bytes(n_genes, n_chunks, n_corr));
    p.stats_off = take((size_t)4 * n_genes * 8);
    p.shift_off = take((size_t)p.gpad * 4);
    p.inv_off = take((size_t)n_genes * 8);
    p.part_off = take((size_t)p.batch * p.nsplit * p.gpad * 8);
    p.maps_off = take((size_t)p.batch * (size_t)(n_rows > 0 ? n_rows : 1) * 4);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

int chr_perm_row_map(int64_t n, uint64_t seed, int draw, int32_t* row_map_out, chr_stream_t stream) {
    using namespace chr;
    CHR_REQUIRE(row_map_out && n > 0 && n < (1LL << 31) && draw >= 0, "chr_perm_row_map: bad arguments");
    perm_row_map_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, (cudaStream_t)stream>>>(row_map_out, (uint32_t)n, seed, (uint32_t)draw);
    CHR_LAUNCH_CHECK();
    return 0;
}

size_t chr_moran_perm_lattice_workspace_bytes(int64_t n_rows, int64_t n_genes, int n_chunks, int64_t n_corr, int permutations) {
    if (n_rows <= 0 || n_genes <= 0 || n_chunks <= 0 || n_corr < 0) return 256;
    return chr::perm_ws(n_rows, n_genes, n_chunks, n_corr, permutations).total + 512;
}

int chr_moran_perm_lattice_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int64_t n_spots, int k, const int32_t* bands,
                               int n_bands, const int32_t* chunks, int n_chunks, const int32_t* corr, int64_t n_corr,
                               const int32_t* absent_rows, int64_t n_absent, const int32_t* pilot_rows, int n_pilot,
                               const int32_t* spot_of_row, const int32_t* row_of_spot, int permutations, uint64_t seed,
                               unsigned flags, double* out_I, int64_t* out_larger, double* out_sum, double* out_sumsq,
                               void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(spot_of_row && row_of_spot && out_I && out_larger && out_sum && out_sumsq && workspace, "chr_moran_perm_lattice_f32: null pointer argument");
    CHR_REQUIRE(permutations > 0, "chr_moran_perm_lattice_f32: permutations must be positive (got %d)", permutations);
    CHR_REQUIRE(!(flags & CHR_FLAG_PRESHIFTED), "chr_moran_perm_lattice_f32: CHR_FLAG_PRESHIFTED is not supported here");
    CHR_REQUIRE(n_rows > 0 && n_rows < (1LL << 31), "chr_moran_perm_lattice_f32: n_rows out of range");
    PermWs p = perm_ws(n_rows, n_genes, n_chunks, n_corr, permutations);
    CHR_REQUIRE(workspace_bytes >= p.total + 256, "chr_moran_perm_lattice_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* stats = (double*)(ws + p.stats_off);
    float* shift = (float*)(ws + p.shift_off);
    double* inv_kzz = (double*)(ws + p.inv_off);
    double* part = (double*)(ws + p.part_off);
    int32_t* maps = (int32_t*)(ws + p.maps_off);
    // the observed statistic (validates the arguments as well) + per-gene mean and z'z
    const int rc = chr_moran_lattice_f32(X, n_rows, n_genes, ld, n_spots, k, bands, n_bands, 0, chunks, n_chunks, corr, n_corr, absent_rows, n_absent,
                                         pilot_rows, n_pilot, flags & ~(CHR_FLAG_GEARY | CHR_FLAG_TIMING), out_I, nullptr, stats, ws + p.lat_off,
                                         chr_moran_lattice_workspace_bytes(n_genes, n_chunks, n_corr), stream);
    if (rc) return rc;
    perm_prepare_kernel<<<(unsigned)ceil_div<int64_t>(p.gpad, 256), 256, 0, st>>>(stats, n_genes, p.gpad, k, shift, inv_kzz,
                                                                                  (long long*)out_larger, out_sum, out_sumsq);
    CHR_LAUNCH_CHECK();
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    const uint32_t rb = (uint32_t)(ld * 4);
    // variant 2: 16-byte cp.async producers (the first design); default: one bulk copy per gathered row
    const bool bulk = ((flags >> CHR_VARIANT_SHIFT) & 0xffu) != 2u && ld % 4 == 0;
    CHR_CUDA(cudaFuncSetAttribute(moran_lattice_perm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
    CHR_CUDA(cudaFuncSetAttribute(moran_lattice_perm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
    const bool timing = flags & CHR_FLAG_TIMING;
    timer_begin(CHR_FAMILY_MORAN, timing, st);
    for (int d0 = 0; d0 < permutations; d0 += p.batch) {
        const int nd = permutations - d0 < p.batch ? permutations - d0 : p.batch;
        perm_maps_kernel<<<dim3((unsigned)ceil_div<int64_t>(n_rows, 256), (unsigned)nd), 256, 0, st>>>(spot_of_row, row_of_spot, n_rows,
                                                                                                      (uint32_t)n_spots, seed, (uint32_t)d0, maps);
        dim3 grid((unsigned)(n_chunks * nd), (unsigned)p.ntile);
        if (bulk)
            moran_lattice_perm_kernel<true><<<grid, kLatThreads, kLatSmem, st>>>(X, rb, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks,
                                                                                 n_chunks, maps, n_rows, shift, part, p.nsplit, p.gpad);
        else
            moran_lattice_perm_kernel<false><<<grid, kLatThreads, kLatSmem, st>>>(X, rb, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks,
                                                                                  n_chunks, maps, n_rows, shift, part, p.nsplit, p.gpad);
        if (p.ncorr_split > 0) {
            dim3 gc((unsigned)p.ntile, (unsigned)p.ncorr_split, (unsigned)nd);
            moran_lattice_perm_corr_kernel<<<gc, 32 * kCorrWarps, 0, st>>>(X, rb, (int)ncol, (const int4*)corr, n_corr, maps, n_rows, shift, part,
                                                                           p.nsplit, n_chunks, p.gpad);
        }
        perm_fold_kernel<<<(unsigned)ceil_div<int64_t>(n_genes, 128), 128, 0, st>>>(part, nd, p.nsplit, p.gpad, n_genes, inv_kzz, out_I,
                                                                                    (long long*)out_larger, out_sum, out_sumsq);
        CHR_LAUNCH_CHECK();
    }
    timer_end(CHR_FAMILY_MORAN, timing, st);
    return 0;
}

}  // extern "C"
