// K3/K4 on an exact square lattice (Visium HD / Stereo-seq bins, k = 8): Moran's I and Geary's C for every gene in one
// streaming pass, with the spatial lag formed in REGISTERS instead of shared-memory gathers.
//
// Replaces the same reference loop as moran.cu (esda.moran.Moran(gene_matrix[c], w).I [+ Geary(...).C],
// /root/reference/chrysalis/core.py:71-76) when the spots sit on a square lattice and W is the 8-neighbour kNN graph
// (core.py:64-65 with neighbors=8, the authors' setting for HD / Stereo-seq: A5:41, A7:94).
//
// Idea.  k z'Wz = z'Az with A the kNN adjacency.  On a lattice A = B + D, where B is the symmetric "queen" stencil
// between PRESENT nodes and D = A - B is a short list of corrections at tissue edges / holes (a boundary bin has fewer
// than 8 stencil neighbours, so kNN reaches further out).  z'Bz needs every unordered stencil pair once; node (x, y) owns
// its pairs with E = (x+1, y), SW = (x-1, y+1), S = (x, y+1), SE = (x+1, y+1):
//        z'Bz = 2 sum_(x,y) z(x,y) [ z(x+1,y) + z(x-1,y+1) + z(x,y+1) + z(x+1,y+1) ].
// Layout ("lattice layout", built by device.lattice_plan): the lattice is cut into BANDS of <= 60 lattice rows; inside a
// band the dense matrix stores column after column (x ascending), each column = the band's rows (y ascending) as
// consecutive matrix rows.  A CTA owns 128 genes and a CHUNK of consecutive columns of one band; its 10 compute warps own
// 6 lattice rows each and walk along x holding column x in registers: one new column (7 rows: 6 own + the row below, one
// 512-byte ld.shared each) per step gives, with a running pair sum, all four forward neighbours of 6 nodes:
// 7/6 shared-memory reads per matrix element (the quad-plan kernel of moran.cu needs ~4.7) and no index lists at all.
// 4 producer warps stream the columns HBM -> shared memory (cp.async, 3 stages x 2 columns, mbarrier full/empty pairs,
// no CTA barrier in the loop; measured at cfg4: 1 / 2 / 4 producer warps = 15.2 / 9.9 / 7.8 ms - the producers' issue rate,
// not the compute warps, sets the pace); the row below a band (first row of the next band) and rows/columns outside the stored
// range are copied from the pilot vector, i.e. they are exactly neutral after the shift.
// Absent nodes inside the stored range hold the pilot value as well (written here, moran_lattice_fill_kernel).
//
// Numerics: as moran.cu - values are shifted by a per-gene pilot mean before any product; fp32 over 8 columns x 6 rows,
// fp32 per thread over the chunk, fp64 across warps / chunks in fixed order; the summation tree depends on the lattice
// only (bit-identical results under gene sharding).
#include "moran_lattice.cuh"

namespace chr {

// ---- pilot shift: mean of a fixed sample of PRESENT rows, per gene (fixed summation order) -------------------------
__global__ void __launch_bounds__(512)
moran_lattice_pilot_kernel(const float* __restrict__ X, int64_t ld, int64_t ncol, const int32_t* __restrict__ rows, int ns,
                           float* __restrict__ pilot) {
    __shared__ float red[8][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + tx;
    float s = 0.f;
    if (col < ncol) {
#pragma unroll 8
        for (int t = ty; t < ns; t += 8) s += __ldg(X + (int64_t)__ldg(rows + t) * ld + col);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) a += red[y][tx];
        pilot[col] = a / (float)ns;
    }
}

// absent nodes of the stored range take the pilot value: exactly neutral after the shift
__global__ void moran_lattice_fill_kernel(float* __restrict__ X, int64_t ld, int64_t ncol4, const int32_t* __restrict__ absent,
                                          int64_t n_absent, const float* __restrict__ pilot) {
    const int64_t c4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 column
    if (c4 >= ncol4) return;
    const float4 p = *reinterpret_cast<const float4*>(pilot + 4 * c4);
    for (int64_t a = blockIdx.y; a < n_absent; a += gridDim.y)
        *reinterpret_cast<float4*>(X + (int64_t)absent[a] * ld + 4 * c4) = p;
}

// one walking step: `nxt` <- column j of the chunk (7 rows of this warp), accumulate the nodes of column j - 1 (`cur`)
template <bool SHIFT, bool MASKED>
__device__ __forceinline__ void lattice_step(uint32_t addr, bool accumulate, const Row4& p, const float (&m)[kLatR], Row4 (&cur)[kLatR + 1],
                                             Row4 (&nxt)[kLatR + 1], Row4 (&pc)[kLatR + 1], Row4 (&l1)[3]) {
#pragma unroll
    for (int r = 0; r <= kLatR; ++r) {
        nxt[r] = lat_lds_row4(addr + r * kLatRowBytes);
        if (SHIFT) r4_sub(nxt[r], p);
    }
    if (accumulate) {
#pragma unroll
        for (int r = 0; r < kLatR; ++r) {
            Row4 s = pc[r + 1];                       // (x-1, y+1) + (x, y+1)
            r4_add(s, nxt[r + 1]);                    // + (x+1, y+1)
            r4_add(s, nxt[r]);                        // + (x+1, y)
            if (MASKED) {                             // short band: rows at / below the halo row do not count
                Row4 mc;
                const f2_t mm = f2_dup(m[r]);
                const f2_t z2 = f2_pack(0.f, 0.f);
                mc.a = f2_fma(mm, cur[r].a, z2);
                mc.b = f2_fma(mm, cur[r].b, z2);
                r4_add(l1[0], mc);
                r4_fma(l1[1], mc, cur[r]);
                r4_fma(l1[2], mc, s);
            } else {
                r4_add(l1[0], cur[r]);
                r4_fma(l1[1], cur[r], cur[r]);
                r4_fma(l1[2], cur[r], s);
            }
        }
    }
#pragma unroll
    for (int r = 1; r <= kLatR; ++r) {                // running pair sum (x, y) + (x+1, y) for the next step
        pc[r] = cur[r];
        r4_add(pc[r], nxt[r]);
    }
}

// bands[b]  = {first matrix row, x of the first stored column, stored columns, lattice rows (<= 60)}
// chunks[c] = {band, first column (relative to the band's first stored column), columns, 0}
template <bool SHIFT>
__global__ void __launch_bounds__(kLatThreads, 1)
moran_lattice_kernel(const float* __restrict__ X, uint32_t row_bytes, int ncol, const int4* __restrict__ bands, int n_bands,
                     const int4* __restrict__ chunks, const float* __restrict__ pilot, double* __restrict__ part, int64_t gpad) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    uint32_t smem0 = (smem_u32(smem_dyn) + 127u) & ~127u;
    asm volatile("" : "+r"(smem0));
    const uint32_t full_bar = smem0 + kLatStages * kLatStageBytes, empty_bar = full_bar + 8 * kLatStages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.x * kGeneTile;
    const int4 ch = __ldg(chunks + blockIdx.y);
    const int4 bd = __ldg(bands + ch.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kLatStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar + 8 * s), "r"(32 * kLatProducers) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_bar + 8 * s), "r"(kLatWarps) : "memory");
        }
        mbar_fence_init();
    }
    __syncthreads();

    int col = tile_col + lane * 4;
    const bool col_ok = col < ncol;
    if (!col_ok) col = tile_col;                                      // idle lanes shadow lane 0 (results unused)
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int col_bytes = col_ok ? 16 : 0;
    const int n_steps = ch.z + 2;                                     // columns j = -1 .. ch.z of the chunk
    const int n_iters = (n_steps + kLatStageCols - 1) / kLatStageCols;

    if (warp >= kLatWarps) {
        // ===== producer warps: stream the chunk's columns, two stages ahead of the compute warps =====
        const int pw = warp - kLatWarps;
        const bool has_next = ch.x + 1 < n_bands;
        int4 nb = make_int4(0, 0, 0, 0);
        if (has_next) nb = __ldg(bands + ch.x + 1);
        const char* pil = reinterpret_cast<const char*>(pilot + col);
        int sidx = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_iters; ++it) {
            if (it >= kLatStages) lat_mbar_wait(empty_bar + 8 * sidx, phase ^ 1u);
            const uint32_t st = smem0 + sidx * kLatStageBytes + lane * 16;
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) {
                const int j = it * kLatStageCols + cc - 1;
                if (j > ch.z) break;
                const int xc = ch.y + j;                              // column relative to the band's first stored column
                const bool valid = xc >= 0 && xc < bd.z;
                const char* src = base + (uint64_t)(uint32_t)(bd.x + (valid ? xc : 0) * bd.w) * row_bytes;
                const char* halo = pil;                               // row below the band: first row of the next band, same x
                if (has_next) {                                       // also for a column outside this band's stored range:
                    const int xn = bd.y + xc - nb.y;                  // its halo node is the SW / SE neighbour of an edge column
                    if (xn >= 0 && xn < nb.z) halo = base + (uint64_t)(uint32_t)(nb.x + xn * nb.w) * row_bytes;
                }
                const int rows_here = valid ? bd.w : 0;
                const uint32_t dst = st + cc * kLatSegBytes;
#pragma unroll 8
                for (int t = pw; t < kLatSegRows; t += kLatProducers) {
                    const char* q = t < rows_here ? src + (uint64_t)t * row_bytes : (t == bd.w ? halo : pil);
                    lat_cp_async_16(dst + t * kLatRowBytes, q, col_bytes);
                }
            }
            lat_cp_async_mbar_arrive(full_bar + 8 * sidx);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
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
    Row4 l1[3], l2[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) l1[s].a = l1[s].b = l2[s].a = l2[s].b = z2;

#ifdef LAT_NO_COMPUTE   /* timing experiment: the producer pipeline alone (wrong statistics) */
    if (warp < kLatWarps) {
        int sidx = 0;
        uint32_t phase = 0;
        for (int it = 0; it < n_iters; ++it) {
            lat_mbar_wait(full_bar + 8 * sidx, phase);
            if (lane == 0) lat_mbar_arrive(empty_bar + 8 * sidx);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
        }
    }
#else
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
            // column j0 into `cb` (accumulating the nodes held in `ca`), then column j0 + 1 into `ca`
            if (masked) {
                lattice_step<SHIFT, true>(a0, j0 != 0, p, m, ca, cb, pc, l1);
                if (j0 + 1 <= ch.z) lattice_step<SHIFT, true>(a0 + kLatSegBytes, j0 + 1 != 0, p, m, cb, ca, pc, l1);
            } else {
                lattice_step<SHIFT, false>(a0, j0 != 0, p, m, ca, cb, pc, l1);
                if (j0 + 1 <= ch.z) lattice_step<SHIFT, false>(a0 + kLatSegBytes, j0 + 1 != 0, p, m, cb, ca, pc, l1);
            }
            __syncwarp();
            if (lane == 0) lat_mbar_arrive(empty_bar + 8 * sidx);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
            if ((it & (kLatFlushStages - 1)) == kLatFlushStages - 1) {
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    r4_add(l2[s], l1[s]);
                    l1[s].a = l1[s].b = z2;
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) r4_add(l2[s], l1[s]);
    }
#endif

    // per-thread fp32 sums -> fp64, cross-warp reduction in fixed order, one statistic at a time
    __syncthreads();                                                  // every column consumed: the stages are free
    double* red = reinterpret_cast<double*>(smem_dyn + (smem0 - smem_u32(smem_dyn)));   // [kLatWarps][kGeneTile]
#pragma unroll
    for (int st = 0; st < kNStat; ++st) {
        if (st < 3 && warp < kLatWarps) {
            float f0, f1, f2, f3;
            f2_unpack(l2[st].a, f0, f1);
            f2_unpack(l2[st].b, f2, f3);
            red[warp * kGeneTile + lane * 4 + 0] = (double)f0;
            red[warp * kGeneTile + lane * 4 + 1] = (double)f1;
            red[warp * kGeneTile + lane * 4 + 2] = (double)f2;
            red[warp * kGeneTile + lane * 4 + 3] = (double)f3;
        }
        __syncthreads();
        if (threadIdx.x < kGeneTile) {
            double t = 0.0;
            if (st < 3) {
#pragma unroll
                for (int w = 0; w < kLatWarps; ++w) t += red[w * kGeneTile + threadIdx.x];
                if (st == 2) t *= 2.0;                                // every unordered stencil pair was visited once
            }
            part[((int64_t)blockIdx.y * kNStat + st) * gpad + tile_col + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// corrections D = A - B and the (indeg - k) weights: entry {row_i, row_j, coefficient (float bits), kind}
//   kind 0:  x'Ax' += c x'_i x'_j          kind 1:  sum (indeg - k) x' += c x'_i,   sum (indeg - k) x'^2 += c x'_i^2
__global__ void __launch_bounds__(32 * kCorrWarps)
moran_lattice_corr_kernel(const float* __restrict__ X, uint32_t row_bytes, int ncol, const int4* __restrict__ corr, int64_t n_corr,
                          const float* __restrict__ pilot, double* __restrict__ part, int64_t gpad, int split0) {
    __shared__ double red[kCorrWarps][kGeneTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.x * kGeneTile;
    int col = tile_col + lane * 4;
    if (col >= ncol) col = tile_col;
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const float4 pf = __ldg(reinterpret_cast<const float4*>(pilot + col));
    const int64_t e0 = (int64_t)blockIdx.y * kCorrPerCta, e1 = e0 + kCorrPerCta < n_corr ? e0 + kCorrPerCta : n_corr;
    double acc[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[s][c] = 0.0;
    // four entries per warp in flight (the two row loads of an entry depend on the entry itself: without this the kernel is a
    // chain of L2 round trips and costs 0.3 ms at cfg4)
    constexpr int kU = 4;
    for (int64_t e = e0 + warp * kU; e < e1; e += kCorrWarps * kU) {
        int4 en[kU];
        float4 xi[kU], xj[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) en[u] = e + u < e1 ? __ldg(corr + e + u) : make_int4(0, 0, 0, 1);       // kind 1, coefficient +0: no effect
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            xi[u] = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)en[u].x * row_bytes));
            xj[u] = __ldg(reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)(en[u].w == 0 ? en[u].y : en[u].x) * row_bytes));
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const float c = __int_as_float(en[u].z);
            const float a[4] = {xi[u].x - pf.x, xi[u].y - pf.y, xi[u].z - pf.z, xi[u].w - pf.w};
            if (en[u].w == 0) {
                const float b[4] = {xj[u].x - pf.x, xj[u].y - pf.y, xj[u].z - pf.z, xj[u].w - pf.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[0][q] += (double)c * (double)a[q] * (double)b[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[1][q] += (double)c * (double)a[q];
                    acc[2][q] += (double)c * (double)a[q] * (double)a[q];
                }
            }
        }
    }
    for (int st = 0; st < kNStat; ++st) {
        if (st >= 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red[warp][lane * 4 + q] = acc[st - 2][q];
        }
        __syncthreads();
        if (threadIdx.x < kGeneTile) {
            double t = 0.0;
            if (st >= 2)
                for (int w = 0; w < kCorrWarps; ++w) t += red[w][threadIdx.x];
            part[((int64_t)(split0 + blockIdx.y) * kNStat + st) * gpad + tile_col + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

struct LatWs {
    int64_t gpad;
    int ntile, nsplit, ncorr_split;
    size_t pilot_off, part_off, total;
};

static LatWs lattice_ws(int64_t ncol, int n_chunks, int64_t n_corr) {
    LatWs p;
    p.ntile = (int)ceil_div<int64_t>(ncol, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    p.ncorr_split = (int)ceil_div<int64_t>(n_corr, kCorrPerCta);
    p.nsplit = n_chunks + p.ncorr_split;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 256); return r; };
    p.pilot_off = take((size_t)p.gpad * sizeof(float));
    p.part_off = take((size_t)p.nsplit * kNStat * p.gpad * sizeof(double));
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_moran_lattice_workspace_bytes(int64_t n_genes, int n_chunks, int64_t n_corr) {
    if (n_genes <= 0 || n_chunks <= 0 || n_corr < 0) return 256;
    return chr::lattice_ws(chr::round_up<int64_t>(n_genes, 4), n_chunks, n_corr).total + 512;
}

int chr_moran_lattice_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int64_t n_spots, int k, const int32_t* bands,
                          int n_bands, const int32_t* chunks, int n_chunks, const int32_t* corr, int64_t n_corr,
                          const int32_t* absent_rows, int64_t n_absent, const int32_t* pilot_rows, int n_pilot, unsigned flags,
                          double* out_I, double* out_C, double* out_stats, void* workspace, size_t workspace_bytes,
                          chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(X && bands && chunks && pilot_rows && out_I && workspace, "chr_moran_lattice_f32: null pointer argument");
    CHR_REQUIRE(n_rows >= n_spots && n_spots > 1 && n_genes > 0, "chr_moran_lattice_f32: need n_rows >= n_spots > 1 and n_genes > 0");
    CHR_REQUIRE(k == 8, "chr_moran_lattice_f32: the stencil path is the 8-neighbour square lattice (k=%d)", k);
    CHR_REQUIRE(ld >= n_genes && ld % 4 == 0 && ((uintptr_t)X & 15) == 0, "chr_moran_lattice_f32: X must be 16-byte aligned with ld %% 4 == 0, ld >= n_genes");
    CHR_REQUIRE(n_rows < (1LL << 31) && ld * 4 < (1LL << 32), "chr_moran_lattice_f32: needs n_rows < 2^31, row bytes < 2^32");
    CHR_REQUIRE(n_bands > 0 && n_chunks > 0 && n_chunks < 65536 && n_pilot > 0 && n_pilot <= kLatPilotMax, "chr_moran_lattice_f32: bad plan sizes");
    CHR_REQUIRE((n_corr == 0 || corr) && (n_absent == 0 || absent_rows) && n_corr >= 0 && n_absent >= 0, "chr_moran_lattice_f32: bad correction / absent lists");
    const bool geary = flags & CHR_FLAG_GEARY, timing = flags & CHR_FLAG_TIMING, preshifted = flags & CHR_FLAG_PRESHIFTED;
    CHR_REQUIRE(!geary || out_C, "chr_moran_lattice_f32: CHR_FLAG_GEARY needs out_C");
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    LatWs p = lattice_ws(ncol, n_chunks, n_corr);
    CHR_REQUIRE(p.ncorr_split < 65536, "chr_moran_lattice_f32: too many corrections (%lld) - use the general kernel", (long long)n_corr);
    CHR_REQUIRE(workspace_bytes >= p.total + 256, "chr_moran_lattice_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* pilot = (float*)(ws + p.pilot_off);
    double* part = (double*)(ws + p.part_off);

    if (preshifted) CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)p.gpad * sizeof(float), st));
    else {
        CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)p.gpad * sizeof(float), st));      // columns ncol .. gpad are read as halo source
        moran_lattice_pilot_kernel<<<(unsigned)ceil_div<int64_t>(ncol, 64), 512, 0, st>>>(X, ld, ncol, pilot_rows, n_pilot, pilot);
    }
    CHR_LAUNCH_CHECK();
    if (n_absent > 0) {
        const int64_t ncol4 = ncol / 4;
        dim3 g((unsigned)ceil_div<int64_t>(ncol4, 128), (unsigned)(n_absent < 1024 ? n_absent : 1024));
        moran_lattice_fill_kernel<<<g, 128, 0, st>>>(X, ld, ncol4, absent_rows, n_absent, pilot);
        CHR_LAUNCH_CHECK();
    }
    const uint32_t rb = (uint32_t)(ld * 4);
    dim3 grid((unsigned)p.ntile, (unsigned)n_chunks);
    timer_begin(CHR_FAMILY_MORAN, timing, st);
    if (preshifted) {
        CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
        moran_lattice_kernel<false><<<grid, kLatThreads, kLatSmem, st>>>(X, rb, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks, pilot, part, p.gpad);
    } else {
        CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
        moran_lattice_kernel<true><<<grid, kLatThreads, kLatSmem, st>>>(X, rb, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks, pilot, part, p.gpad);
    }
    timer_end(CHR_FAMILY_MORAN, timing, st);
    CHR_LAUNCH_CHECK();
    if (p.ncorr_split > 0) {
        dim3 gc((unsigned)p.ntile, (unsigned)p.ncorr_split);
        moran_lattice_corr_kernel<<<gc, 32 * kCorrWarps, 0, st>>>(X, rb, (int)ncol, (const int4*)corr, n_corr, pilot, part, p.gpad, n_chunks);
        CHR_LAUNCH_CHECK();
    }
    return launch_moran_finalize(part, p.nsplit, p.gpad, n_genes, n_spots, k, pilot, geary, out_I, out_C, out_stats, st);
}

}  // extern "C"
