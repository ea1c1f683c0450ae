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
//
// Where the columns come from (LatSrc):
//   dense  X [rows][ld] float32 in lattice layout: 4 producer warps copy rows with cp.async (HBM-bound: one read of X)
//   CSR    the counts themselves, prepared once per call by lattice_csr_prepare_*: per stored entry the position inside
//          its 128-gene slab (ecol, 0xFF = gene filtered out) and log1p(count * scale) (eval); per (slab, stored row) the
//          range of the row's entries that fall into the slab (eoff).  6 producer warps, one staged column each, build the
//          tiles in shared memory: rows are initialised (zeros for a present node, the pilot vector for an absent one) and
//          the ~6 entries a row has inside the slab are scattered in (8 loads per lane in flight, the entry ranges of a warp's
//          next column prefetched).  No dense matrix exists: HBM sees the entries once (5 bytes each) instead of 4 bytes
//          per spot and gene.
struct LatSrc {
    const float* X;                 // dense
    uint32_t row_bytes;
    const uint32_t* eoff;           // CSR: [n_rows][n_slabs + 1] absolute entry offsets, 0xFFFFFFFF in every slot of an absent node
    const uint8_t* ecol;
    const float* eval;
    int64_t n_slabs1;               // n_slabs + 1
};

constexpr int kLatCsrProducers = kLatStages * kLatStageCols;      // 6: one producer warp per staged column
constexpr int kLatCsrThreads = 32 * (kLatWarps + kLatCsrProducers);
constexpr int kLatTmaThreads = 32 * (kLatWarps + 1);              // one producer warp: TMA boxes need no per-row instructions

// the three boxes of the TMA producer over X [rows][ld]: a whole column of a full band (60 rows x 512 B), a column of the
// short last band, a single row (the row below a band)
struct LatMaps {
    CUtensorMap full, last, row;
};

enum { kProdCpAsync = 0, kProdCsr = 1, kProdTma = 2 };

template <bool SHIFT, int PROD>
__global__ void __launch_bounds__(PROD == kProdCsr ? kLatCsrThreads : (PROD == kProdTma ? kLatTmaThreads : kLatThreads), 1)
moran_lattice_kernel(const LatSrc src, const __grid_constant__ LatMaps maps, int ncol, const int4* __restrict__ bands, int n_bands,
                     const int4* __restrict__ chunks, const float* __restrict__ pilot, double* __restrict__ part, int64_t gpad) {
    constexpr bool CSR = PROD == kProdCsr, TMA = PROD == kProdTma;
    constexpr int kProducers = CSR ? kLatCsrProducers : (TMA ? 1 : kLatProducers);
    const float* __restrict__ X = src.X;
    const uint32_t row_bytes = src.row_bytes;
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
            // cp.async: every producer thread arrives through its cp.async group; CSR: the 2 warps that build the stage's columns
            // arrive; TMA: one expect_tx arrival + the producer warp's 32 cp.async arrivals (rows the boxes cannot deliver)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_bar + 8 * s), "r"(CSR ? 32 * kLatStageCols : (TMA ? 33 : 32 * kProducers)) : "memory");
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

    if (CSR && warp >= kLatWarps) {
        // ===== CSR producers: warp p owns the staged column slot (stage p / 2, column p % 2) for the whole chunk =====
        const int pw = warp - kLatWarps;
        const int my_stage = pw / kLatStageCols, cc = pw % kLatStageCols;
        const bool has_next = ch.x + 1 < n_bands;
        int4 nb = make_int4(0, 0, 0, 0);
        if (has_next) nb = __ldg(bands + ch.x + 1);
        const float4 pil4 = __ldg(reinterpret_cast<const float4*>(pilot + col));
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t* __restrict__ off = src.eoff + blockIdx.x;       // eoff[row][slab], eoff[row][slab + 1]
        const int64_t off_ld = src.n_slabs1;
        const uint32_t seg = smem0 + my_stage * kLatStageBytes + cc * kLatSegBytes;
        // entry ranges of rows `lane` and `lane + 32` of column j (0xFFFFFFFF = absent node / nothing stored there)
        auto resolve = [&](int j, uint2 (&rng)[2]) {
            const int xc = ch.y + j;
            const bool valid = xc >= 0 && xc < bd.z;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = lane + 32 * h;
                int r = -1;
                if (j <= ch.z) {
                    if (t < bd.w) {
                        if (valid) r = bd.x + xc * bd.w + t;
                    } else if (t == bd.w && has_next) {
                        const int xn = bd.y + xc - nb.y;
                        if (xn >= 0 && xn < nb.z) r = nb.x + xn * nb.w;
                    }
                }
                rng[h] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                if (r >= 0) rng[h] = make_uint2(__ldg(off + (int64_t)r * off_ld), __ldg(off + (int64_t)r * off_ld + 1));
            }
        };
        uint2 cur[2], nxt[2];
        resolve(my_stage * kLatStageCols + cc - 1, cur);
        uint32_t phase = 0;
#pragma unroll 1
        for (int it = my_stage; it < n_iters; it += kLatStages, phase ^= 1u) {
            const int j = it * kLatStageCols + cc - 1;
            resolve((it + kLatStages) * kLatStageCols + cc - 1, nxt);   // the ranges of this warp's next column are in flight meanwhile
            // first batch of entries of both rows of this lane: loaded into registers BEFORE the stage is waited for and the
            // rows are initialised, so their latency hides behind both
            constexpr int kB = 12;
            unsigned c0[kB], c1[kB];
            float v0[kB], v1[kB];
            const uint32_t b0 = cur[0].x, b1 = cur[1].x;                 // absent node (0xFFFFFFFF): no entries
            const uint32_t n0 = b0 == 0xFFFFFFFFu ? 0u : cur[0].y - b0, n1 = b1 == 0xFFFFFFFFu ? 0u : cur[1].y - b1;
#pragma unroll
            for (int q = 0; q < kB; ++q) {
                const bool in0 = (uint32_t)q < n0, in1 = (uint32_t)q < n1;
                c0[q] = in0 ? (unsigned)__ldg(src.ecol + b0 + q) : 0xFFu;
                v0[q] = in0 ? __ldg(src.eval + b0 + q) : 0.f;
                c1[q] = in1 ? (unsigned)__ldg(src.ecol + b1 + q) : 0xFFu;
                v1[q] = in1 ? __ldg(src.eval + b1 + q) : 0.f;
            }
            if (it >= kLatStages) lat_mbar_wait(empty_bar + 8 * my_stage, phase ^ 1u);
            if (j <= ch.z) {
                const unsigned pm0 = __ballot_sync(0xffffffffu, cur[0].x != 0xFFFFFFFFu), pm1 = __ballot_sync(0xffffffffu, cur[1].x != 0xFFFFFFFFu);
                // rows start as zeros (present node: no entry = no expression) or as the pilot vector (absent node: neutral)
#pragma unroll 8
                for (int t = 0; t < kLatSegRows; ++t) {
                    const bool pres = ((t < 32 ? pm0 : pm1) >> (t & 31)) & 1u;
                    const float4 v = pres ? zero4 : pil4;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(seg + t * kLatRowBytes + lane * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                __syncwarp();
                const uint32_t row0 = seg + (uint32_t)lane * kLatRowBytes, row1 = row0 + 32 * kLatRowBytes;
#pragma unroll
                for (int q = 0; q < kB; ++q) {
                    if (c0[q] != 0xFFu) asm volatile("st.shared.f32 [%0], %1;" ::"r"(row0 + c0[q] * 4), "f"(v0[q]) : "memory");
                    if (c1[q] != 0xFFu) asm volatile("st.shared.f32 [%0], %1;" ::"r"(row1 + c1[q] * 4), "f"(v1[q]) : "memory");
                }
                // rows with more than kB entries inside the slab (rare)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t row_addr = h ? row1 : row0, bb = h ? b1 : b0, nn = h ? n1 : n0;
                    for (uint32_t o = kB; o < nn; o += 8) {
                        unsigned c[8];
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const bool in = o + q < nn;
                            c[q] = in ? (unsigned)__ldg(src.ecol + bb + o + q) : 0xFFu;
                            v[q] = in ? __ldg(src.eval + bb + o + q) : 0.f;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (c[q] != 0xFFu) asm volatile("st.shared.f32 [%0], %1;" ::"r"(row_addr + c[q] * 4), "f"(v[q]) : "memory");
                    }
                }
            }
            __syncwarp();
            lat_mbar_arrive(full_bar + 8 * my_stage);                 // release: the stores above are visible to the waiters
            cur[0] = nxt[0];
            cur[1] = nxt[1];
        }
    }

    if (TMA && warp >= kLatWarps) {
        // ===== TMA producer: one box per column (cp.async.bulk.tensor.2d, 60 rows x 512 B), one box for the row below =====
        // Rows no box can deliver - the pilot vector for columns / halo rows outside the stored range and for the rows past the
        // end of a short band - are copied by the same warp with cp.async; both kinds of completion land on the stage's barrier.
        const bool has_next = ch.x + 1 < n_bands;
        int4 nb = make_int4(0, 0, 0, 0);
        if (has_next) nb = __ldg(bands + ch.x + 1);
        const char* pil = reinterpret_cast<const char*>(pilot + col);
        const CUtensorMap* col_map = bd.w == kLatBH ? &maps.full : &maps.last;
        if (lane == 0) { prefetch_tensormap(col_map); prefetch_tensormap(&maps.row); }
        int sidx = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_iters; ++it) {
            if (it >= kLatStages) lat_mbar_wait(empty_bar + 8 * sidx, phase ^ 1u);
            const uint32_t st = smem0 + sidx * kLatStageBytes;
            const uint32_t bar = full_bar + 8 * sidx;
            // bytes the boxes of this stage will deliver
            uint32_t tx = 0;
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) {
                const int j = it * kLatStageCols + cc - 1;
                if (j > ch.z) break;
                const int xc = ch.y + j;
                if (xc >= 0 && xc < bd.z) tx += (uint32_t)bd.w * kLatRowBytes;
                if (has_next) {
                    const int xn = bd.y + xc - nb.y;
                    if (xn >= 0 && xn < nb.z) tx += kLatRowBytes;
                }
            }
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
#pragma unroll
            for (int cc = 0; cc < kLatStageCols; ++cc) {
                const int j = it * kLatStageCols + cc - 1;
                if (j > ch.z) break;
                const int xc = ch.y + j;
                const bool valid = xc >= 0 && xc < bd.z;
                const uint32_t dst = st + cc * kLatSegBytes;
                int halo_row = -1;
                if (has_next) {
                    const int xn = bd.y + xc - nb.y;
                    if (xn >= 0 && xn < nb.z) halo_row = nb.x + xn * nb.w;
                }
                if (lane == 0) {
                    if (valid)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                                     "l"(col_map), "r"(bar), "r"(tile_col), "r"(bd.x + xc * bd.w) : "memory");
                    if (halo_row >= 0)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst + bd.w * kLatRowBytes),
                                     "l"(&maps.row), "r"(bar), "r"(tile_col), "r"(halo_row) : "memory");
                }
                // the rest of the column from the pilot vector
                const int first_pilot = valid ? (halo_row >= 0 ? bd.w + 1 : bd.w) : 0;
                for (int t = first_pilot; t < kLatSegRows; ++t) {
                    if (!valid && t == bd.w && halo_row >= 0) continue;   // delivered by the row box
                    lat_cp_async_16(dst + t * kLatRowBytes + lane * 16, pil, col_bytes);
                }
            }
            lat_cp_async_mbar_arrive(bar);
            if (++sidx == kLatStages) { sidx = 0; phase ^= 1u; }
        }
    }

    if (PROD == kProdCpAsync && warp >= kLatWarps) {
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

// tensor map over X [rows][ld] float32 with a box of 128 genes x box_rows rows (no swizzle: a row of the box is the 512-byte
// segment a warp reads with one ld.shared)
static int make_lattice_map(CUtensorMap* m, const float* X, int64_t n_rows, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    CHR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not found");
    CHR_REQUIRE(box_rows >= 1 && box_rows <= 256, "lattice tensor map: bad box height %d", box_rows);
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGeneTile, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CHR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for the lattice matrix", (int)r);
    return 0;
}

// ---- CSR-fed variant: preparation, pilot, corrections ---------------------------------------------------------------
// per stored entry: position inside its 128-gene slab (0xFF: gene filtered out) and the value the dense matrix would hold
// (same arithmetic as csr_densify_kernel: product in float64, rounded once, log1pf)
__global__ void lattice_csr_entries_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
                                           int64_t n, const int32_t* __restrict__ gene_col, const double* __restrict__ scale, int apply_log1p,
                                           uint8_t* __restrict__ ecol, float* __restrict__ eval, int32_t* __restrict__ unsorted) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t b = indptr[i], e = indptr[i + 1];
        const double sc = scale ? scale[i] : 1.0;
        for (int64_t t = b + lane; t < e; t += 32) {
            const int32_t g = indices[t];
            if (unsorted && t > b && indices[t - 1] >= g) *unsorted = 1;      // the slab ranges need strictly increasing columns per row
            const int c = gene_col[g];
            float v = scale ? (float)((double)data[t] * sc) : data[t];
            if (apply_log1p) v = log1pf(v);
            ecol[t] = c < 0 ? (uint8_t)0xFF : (uint8_t)(c & (kGeneTile - 1));
            eval[t] = v;
        }
    }
}

// slab_first[s] = original column of the first kept gene of slab s; slab_first[n_slabs] = n_genes_all
__global__ void lattice_csr_slab_first_kernel(const int32_t* __restrict__ gene_col, int64_t n_genes_all, int n_slabs, int32_t* __restrict__ slab_first) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) slab_first[n_slabs] = (int32_t)n_genes_all;
    if (g >= n_genes_all) return;
    const int c = gene_col[g];
    if (c >= 0 && (c & (kGeneTile - 1)) == 0) slab_first[c / kGeneTile] = (int32_t)g;
}

// eoff[r][s] = absolute index of the first entry of the spot stored at row r whose (original, sorted) column is
// >= slab_first[s]; every slot of an absent node is 0xFFFFFFFF.  One warp per row: the row's columns are staged in
// shared memory and the lanes binary-search the slab boundaries there.
constexpr int kOffWarps = 8, kOffStage = 1024;
__global__ void __launch_bounds__(32 * kOffWarps)
lattice_csr_offsets_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int32_t* __restrict__ spot_of_row,
                           int64_t n_rows, const int32_t* __restrict__ slab_first, int n_slabs1, uint32_t* __restrict__ eoff) {
    __shared__ int32_t cols_s[kOffWarps][kOffStage];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * kOffWarps + warp;
    if (r >= n_rows) return;
    uint32_t* __restrict__ out = eoff + r * n_slabs1;
    const int spot = spot_of_row[r];
    if (spot < 0) {
        for (int s = lane; s < n_slabs1; s += 32) out[s] = 0xFFFFFFFFu;
        return;
    }
    const int64_t b = indptr[spot], e = indptr[spot + 1];
    const int len = (int)(e - b);
    const bool staged = len <= kOffStage;
    if (staged) {
        for (int t = lane; t < len; t += 32) cols_s[warp][t] = __ldg(indices + b + t);
        __syncwarp();
    }
    for (int s = lane; s < n_slabs1; s += 32) {
        const int32_t key = slab_first[s];
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int32_t c = staged ? cols_s[warp][mid] : __ldg(indices + b + mid);
            if (c < key) lo = mid + 1; else hi = mid;
        }
        out[s] = (uint32_t)(b + lo);
    }
}

// pilot shift from the CSR form: mean over the sample rows; one CTA per slab, warp w adds rows w, w + 16, ... into its own
// accumulator, the accumulators are then added in warp order (fixed summation tree)
constexpr int kPilotWarps = 16;
__global__ void __launch_bounds__(32 * kPilotWarps)
lattice_csr_pilot_kernel(const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ ecol, const float* __restrict__ eval, int64_t n_slabs1,
                         const int32_t* __restrict__ rows, int ns, float* __restrict__ pilot) {
    __shared__ float acc[kPilotWarps][kGeneTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = lane; c < kGeneTile; c += 32) acc[warp][c] = 0.f;
    __syncwarp();
    for (int t = warp; t < ns; t += kPilotWarps) {
        const uint32_t* __restrict__ o = eoff + (int64_t)rows[t] * n_slabs1 + blockIdx.x;
        const uint32_t b = o[0], e = o[1];
        for (uint32_t u = b + lane; u < e; u += 32) {
            const unsigned c = ecol[u];
            if (c != 0xFFu) acc[warp][c] += eval[u];                  // one entry per gene and row: no two lanes share a slot
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < kGeneTile) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kPilotWarps; ++w) a += acc[w][threadIdx.x];
        pilot[(int64_t)blockIdx.x * kGeneTile + threadIdx.x] = a / (float)ns;
    }
}

// corrections from the CSR form: the two rows of an entry are rebuilt (128 genes of the slab) in per-warp shared memory
__global__ void __launch_bounds__(32 * kCorrWarps)
moran_lattice_csr_corr_kernel(const uint32_t* __restrict__ eoff, const uint8_t* __restrict__ ecol, const float* __restrict__ eval, int64_t n_slabs1,
                              const int4* __restrict__ corr, int64_t n_corr, const float* __restrict__ pilot, double* __restrict__ part,
                              int64_t gpad, int split0) {
    __shared__ double red[kCorrWarps][kGeneTile];
    __shared__ __align__(16) float rows_s[kCorrWarps][2][kGeneTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_col = blockIdx.x * kGeneTile;
    const float4 pf = __ldg(reinterpret_cast<const float4*>(pilot + tile_col + lane * 4));
    const int64_t e0 = (int64_t)blockIdx.y * kCorrPerCta, e1 = e0 + kCorrPerCta < n_corr ? e0 + kCorrPerCta : n_corr;
    double acc[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[s][c] = 0.0;
    for (int64_t e = e0 + warp; e < e1; e += kCorrWarps) {
        const int4 en = __ldg(corr + e);
        const int nrow = en.w == 0 ? 2 : 1;
        for (int h = 0; h < nrow; ++h) {
            const int r = h == 0 ? en.x : en.y;
            *reinterpret_cast<float4*>(&rows_s[warp][h][lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            const uint32_t* __restrict__ o = eoff + (int64_t)r * n_slabs1 + blockIdx.x;
            for (uint32_t u = o[0] + lane; u < o[1]; u += 32) {
                const unsigned c = ecol[u];
                if (c != 0xFFu) rows_s[warp][h][c] = eval[u];
            }
        }
        __syncwarp();
        const float c = __int_as_float(en.z);
        const float4 xi = *reinterpret_cast<const float4*>(&rows_s[warp][0][lane * 4]);
        const float a[4] = {xi.x - pf.x, xi.y - pf.y, xi.z - pf.z, xi.w - pf.w};
        if (en.w == 0) {
            const float4 xj = *reinterpret_cast<const float4*>(&rows_s[warp][1][lane * 4]);
            const float b[4] = {xj.x - pf.x, xj.y - pf.y, xj.z - pf.z, xj.w - pf.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[0][q] += (double)c * (double)a[q] * (double)b[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[1][q] += (double)c * (double)a[q];
                acc[2][q] += (double)c * (double)a[q] * (double)a[q];
            }
        }
        __syncwarp();
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

struct LatCsrWs {
    LatWs lat;
    int n_slabs;
    size_t lat_off, ecol_off, eval_off, eoff_off, first_off, total;
};

static LatCsrWs lattice_csr_ws(int64_t nnz, int64_t n_rows, int64_t n_keep, int n_chunks, int64_t n_corr) {
    LatCsrWs p;
    const int64_t ncol = round_up<int64_t>(n_keep, 4);
    p.lat = lattice_ws(ncol, n_chunks, n_corr);
    p.n_slabs = p.lat.ntile;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 256); return r; };
    p.lat_off = take(p.lat.total);
    p.ecol_off = take((size_t)(nnz > 0 ? nnz : 1));
    p.eval_off = take((size_t)(nnz > 0 ? nnz : 1) * 4);
    p.eoff_off = take((size_t)(p.n_slabs + 1) * (size_t)n_rows * 4);
    p.first_off = take((size_t)(p.n_slabs + 1) * 4);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_moran_lattice_csr_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_keep, int n_chunks, int64_t n_corr) {
    if (nnz < 0 || n_rows <= 0 || n_keep <= 0 || n_chunks <= 0 || n_corr < 0) return 256;
    return chr::lattice_csr_ws(nnz, n_rows, n_keep, n_chunks, n_corr).total + 512;
}

int chr_moran_lattice_csr_f32(const int64_t* indptr, const int32_t* indices, const float* data, int64_t nnz, int64_t n_spots,
                              int64_t n_genes_all, const int32_t* gene_col, int64_t n_keep, const double* scale, int apply_log1p,
                              int64_t n_rows, int k, const int32_t* bands, int n_bands, const int32_t* chunks, int n_chunks,
                              const int32_t* corr, int64_t n_corr, const int32_t* spot_of_row, const int32_t* pilot_rows, int n_pilot,
                              unsigned flags, double* out_I, double* out_C, double* out_stats, int32_t* unsorted_flag, void* workspace,
                              size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(indptr && indices && data && gene_col && bands && chunks && spot_of_row && pilot_rows && out_I && workspace,
                "chr_moran_lattice_csr_f32: null pointer argument");
    CHR_REQUIRE(n_rows >= n_spots && n_spots > 1 && n_keep > 0 && n_genes_all >= n_keep, "chr_moran_lattice_csr_f32: bad sizes");
    CHR_REQUIRE(nnz >= 0 && nnz < 0xFFFFFFFFLL, "chr_moran_lattice_csr_f32: needs fewer than 2^32 - 1 stored entries (got %lld)", (long long)nnz);
    CHR_REQUIRE(k == 8, "chr_moran_lattice_csr_f32: the stencil path is the 8-neighbour square lattice (k=%d)", k);
    CHR_REQUIRE(n_rows < (1LL << 31) && n_bands > 0 && n_chunks > 0 && n_chunks < 65536 && n_pilot > 0 && n_pilot <= kLatPilotMax,
                "chr_moran_lattice_csr_f32: bad plan sizes");
    CHR_REQUIRE((n_corr == 0 || corr) && n_corr >= 0, "chr_moran_lattice_csr_f32: bad correction list");
    const bool geary = flags & CHR_FLAG_GEARY, timing = flags & CHR_FLAG_TIMING;
    CHR_REQUIRE(!geary || out_C, "chr_moran_lattice_csr_f32: CHR_FLAG_GEARY needs out_C");
    CHR_REQUIRE(!(flags & CHR_FLAG_PRESHIFTED), "chr_moran_lattice_csr_f32: CHR_FLAG_PRESHIFTED has no meaning for sparse input");
    LatCsrWs p = lattice_csr_ws(nnz, n_rows, n_keep, n_chunks, n_corr);
    CHR_REQUIRE(p.lat.ncorr_split < 65536 && p.n_slabs + 1 < 65536, "chr_moran_lattice_csr_f32: too many corrections / gene slabs");
    CHR_REQUIRE(workspace_bytes >= p.total + 256, "chr_moran_lattice_csr_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    char* lws = ws + p.lat_off;
    float* pilot = (float*)(lws + p.lat.pilot_off);
    double* part = (double*)(lws + p.lat.part_off);
    uint8_t* ecol = (uint8_t*)(ws + p.ecol_off);
    float* eval = (float*)(ws + p.eval_off);
    uint32_t* eoff = (uint32_t*)(ws + p.eoff_off);
    int32_t* slab_first = (int32_t*)(ws + p.first_off);

    // preparation: entries -> (position in slab, value); per (slab, stored row) entry ranges
    if (unsorted_flag) CHR_CUDA(cudaMemsetAsync(unsorted_flag, 0, sizeof(int32_t), st));
    lattice_csr_entries_kernel<<<kNumSMs * 8, 256, 0, st>>>(indptr, indices, data, n_spots, gene_col, scale, apply_log1p, ecol, eval, unsorted_flag);
    lattice_csr_slab_first_kernel<<<(unsigned)ceil_div<int64_t>(n_genes_all, 256), 256, 0, st>>>(gene_col, n_genes_all, p.n_slabs, slab_first);
    lattice_csr_offsets_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, kOffWarps), 32 * kOffWarps, 0, st>>>(indptr, indices, spot_of_row, n_rows,
                                                                                                     slab_first, p.n_slabs + 1, eoff);
    lattice_csr_pilot_kernel<<<p.n_slabs, 32 * kPilotWarps, 0, st>>>(eoff, ecol, eval, p.n_slabs + 1, pilot_rows, n_pilot, pilot);
    CHR_LAUNCH_CHECK();

    const int64_t ncol = round_up<int64_t>(n_keep, 4);
    LatSrc src{};
    src.eoff = eoff;
    src.ecol = ecol;
    src.eval = eval;
    src.n_slabs1 = p.n_slabs + 1;
    dim3 grid((unsigned)p.lat.ntile, (unsigned)n_chunks);
    CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<true, kProdCsr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
    LatMaps maps{};
    timer_begin(CHR_FAMILY_MORAN, timing, st);
    moran_lattice_kernel<true, kProdCsr><<<grid, kLatCsrThreads, kLatSmem, st>>>(src, maps, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks,
                                                                               pilot, part, p.lat.gpad);
    timer_end(CHR_FAMILY_MORAN, timing, st);
    CHR_LAUNCH_CHECK();
    if (p.lat.ncorr_split > 0) {
        dim3 gc((unsigned)p.lat.ntile, (unsigned)p.lat.ncorr_split);
        moran_lattice_csr_corr_kernel<<<gc, 32 * kCorrWarps, 0, st>>>(eoff, ecol, eval, p.n_slabs + 1, (const int4*)corr, n_corr, pilot, part,
                                                                      p.lat.gpad, n_chunks);
        CHR_LAUNCH_CHECK();
    }
    return launch_moran_finalize(part, p.lat.nsplit, p.lat.gpad, n_keep, n_spots, k, pilot, geary, out_I, out_C, out_stats, st);
}

size_t chr_moran_lattice_workspace_bytes(int64_t n_genes, int n_chunks, int64_t n_corr) {
    if (n_genes <= 0 || n_chunks <= 0 || n_corr < 0) return 256;
    return chr::lattice_ws(chr::round_up<int64_t>(n_genes, 4), n_chunks, n_corr).total + 512;
}

int chr_moran_lattice_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int64_t n_spots, int k, const int32_t* bands,
                          int n_bands, int last_band_rows, const int32_t* chunks, int n_chunks, const int32_t* corr, int64_t n_corr,
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
    LatSrc src{};
    src.X = X;
    src.row_bytes = rb;
    // default: TMA producer (one tensor-map box per staged column; needs the height of the last band, the only one that may
    // be shorter than 60 rows); kernel variant 2 or last_band_rows == 0: the cp.async producer warps
    CHR_REQUIRE(last_band_rows >= 0 && last_band_rows <= kLatBH, "chr_moran_lattice_f32: last_band_rows out of range (%d)", last_band_rows);
    const bool use_tma = last_band_rows > 0 && ((flags >> CHR_VARIANT_SHIFT) & 0xffu) != 2u && !preshifted;
    LatMaps maps{};
    if (use_tma) {
        if (int rc = make_lattice_map(&maps.full, X, n_rows, ld, kLatBH)) return rc;
        if (int rc = make_lattice_map(&maps.last, X, n_rows, ld, last_band_rows)) return rc;
        if (int rc = make_lattice_map(&maps.row, X, n_rows, ld, 1)) return rc;
    }
    timer_begin(CHR_FAMILY_MORAN, timing, st);
    if (preshifted) {
        CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<false, kProdCpAsync>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
        moran_lattice_kernel<false, kProdCpAsync><<<grid, kLatThreads, kLatSmem, st>>>(src, maps, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks, pilot, part, p.gpad);
    } else {
        if (use_tma) {
            CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<true, kProdTma>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
            moran_lattice_kernel<true, kProdTma><<<grid, kLatTmaThreads, kLatSmem, st>>>(src, maps, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks, pilot, part, p.gpad);
        } else {
            CHR_CUDA(cudaFuncSetAttribute(moran_lattice_kernel<true, kProdCpAsync>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatSmem));
            moran_lattice_kernel<true, kProdCpAsync><<<grid, kLatThreads, kLatSmem, st>>>(src, maps, (int)ncol, (const int4*)bands, n_bands, (const int4*)chunks, pilot, part, p.gpad);
        }
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
