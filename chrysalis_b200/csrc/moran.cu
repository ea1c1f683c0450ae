// K3/K4 - Moran's I (and Geary's C) for every gene of a dense spots x genes matrix, one pass.
//
// Replaces the per-gene loop  esda.moran.Moran(gene_matrix[c], w, permutations=0).I
// [+ esda.geary.Geary(...).C]  of /root/reference/chrysalis/core.py:71-76.
//
// Layout: X is [n_spots][ld] float32 row-major (row = spot, spots in a spatially coherent
// order), W is an ELL kNN list nbr[n_spots][k] with implicit weights 1/k (libpysal
// transform 'R').  A warp owns 128 consecutive genes (one float4 per lane) of the spots it
// visits: the own row and the neighbour rows are coalesced 512-byte segments, every lane
// accumulates its four genes over the spots of the CTA's spot range, so there is no
// cross-thread reduction in the streaming loop.  Re-reads of neighbour rows are served by
// L1/L2 (spots are Hilbert-ordered by the host), HBM sees X once.
//
// Numerics: values are shifted by a per-gene pilot mean p (mean of a fixed
// spot sample) before any product, so the one-pass identities
//     z'z  = S2 - n m'^2,      k z'Wz = SW - m' (k S1 + SS) + m'^2 n k       (m' = S1/n)
// with S1 = sum x', S2 = sum x'^2, SW = sum_i x'_i s_i, SS = sum_i s_i, s_i = sum_{j in N(i)} x'_j
// carry no cancellation; partial sums are fp32 over 16 spots per warp, then fp32 over the CTA's
// 2048 spots, then fp64.
#include "common.cuh"

namespace chr {

constexpr int kMoranThreads = 256;
constexpr int kMoranWarps = kMoranThreads / 32;
constexpr int kGeneTile = 128;   // genes per CTA = 32 lanes x float4
constexpr int kChunk = 16;       // spots per warp between fp32 -> fp64 flushes
constexpr int kNStat = 5;        // S1, S2, SW, SS, SQ (Geary numerator * k)
constexpr int kPilotSample = 512;

// ---- pilot shift: mean of a fixed sample of spots, per gene --------------------------------
__global__ void moran_pilot_kernel(const float* __restrict__ X, int64_t n, int64_t ld, int64_t ncol,
                                   float* __restrict__ pilot) {
    int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int ns = n < kPilotSample ? (int)n : kPilotSample;
    float s = 0.f;
    for (int t = 0; t < ns; ++t) {
        int64_t i = (int64_t)t * n / ns;
        s += __ldg(X + i * ld + col);
    }
    pilot[col] = s / (float)ns;
}

// one elected lane asks the TMA unit to pull a row segment into L2 ahead of use
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// =============================================================================================
// v2 streaming kernel.  Same mathematics, restructured after the first ncu capture
// (profiles/r1a: 243 instructions and ~1700 cycles per spot-row, issue 57 %, 16 warps/SM):
//   * W's neighbour lists are staged in shared memory, 128 spots at a time, by cp.async one
//     chunk ahead, so the gather addresses never wait on a global load
//   * every warp works on U = 2 spots at once: 18 independent 512-byte row loads in flight
//   * all arithmetic is packed f32x2 (FADD2 / FFMA2): half the issue slots of scalar FP32
//   * accumulation is two-level fp32 in registers (16 spots, then 16 x 16 spots); a CTA covers
//     only 2048 spots, partial sums leave the CTA as fp64 - no fp64 and no F2F in the loop
//   * row addresses are one IMAD.WIDE.U32 each (row index x row bytes + base)
// =============================================================================================
typedef unsigned long long f2_t;  // two packed fp32 values in a 64-bit register pair

__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t c; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) { f2_t c; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_pack(float x, float y) { f2_t c; asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(x), "f"(y)); return c; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }

struct Row4 {  // four consecutive genes of one spot
    f2_t a, b;
};
__device__ __forceinline__ Row4 ld_row4(const char* p) {
    Row4 r;
    asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(r.a), "=l"(r.b) : "l"(p));
    return r;
}
__device__ __forceinline__ void r4_add(Row4& s, const Row4& v) { s.a = f2_add(s.a, v.a); s.b = f2_add(s.b, v.b); }
__device__ __forceinline__ void r4_sub(Row4& s, const Row4& v) { s.a = f2_sub(s.a, v.a); s.b = f2_sub(s.b, v.b); }
__device__ __forceinline__ void r4_fma(Row4& s, const Row4& x, const Row4& y) { s.a = f2_fma(x.a, y.a, s.a); s.b = f2_fma(x.b, y.b, s.b); }
__device__ __forceinline__ void r4_sqdiff(Row4& q, const Row4& x, const Row4& v) {
    Row4 d = x;
    r4_sub(d, v);
    r4_fma(q, d, d);
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
// v4 symmetric quad kernel (default).  The r1b capture showed the v2 kernel pinned on the L1 data
// pipe (72 % of its wavefront peak: 9 x 512-byte row reads per spot and 97 instructions per
// row slab).  z'Az only needs every UNORDERED pair {i, j} once:
//        z'Az = sum_{i<j} (a_ij + a_ji) z_i z_j + sum_i a_ii z_i^2
// so W is re-expressed, once per W, as a plan over QUADS of 4 consecutive spots (consecutive
// spots of the Hilbert order are spatial neighbours):
//   * pairs inside a quad      -> 6 coefficients, served from the registers that hold the 4 rows
//   * every other pair         -> owned by exactly ONE endpoint: (row, coefficient) in that
//                                 spot's list; owners are balanced so the 4 lists of a quad have
//                                 nearly equal length R (the quad is processed in R rounds)
// which needs ~3.9 row reads per spot instead of 9 (square lattice, k = 8).  The mean shift uses
// csum_i = sum of the external coefficients of spot i; the cross term sum_j indeg_j x_j replaces
// sum_i s_i (they are the same number), and Geary's numerator follows from one more sum:
//   sum_ij a_ij (x_i - x_j)^2 = k sum x_i^2 + sum_j indeg_j x_j^2 - 2 x'Ax.
// Any neighbour list is legal (duplicates, self loops, asymmetric kNN): coefficients are counts.
// =============================================================================================
constexpr int kSymHdrWords = 16;      // R, nvalid, c01 c02 | c03 c12 c13 c23 | csum[4] | indeg[4]
constexpr int kSymRoundWords = 8;     // row[4] | coef[4]
constexpr int kSymChunkQuads = 32;    // quads staged at once (128 spots)
constexpr int kSymQuadsPerWarp = kSymChunkQuads / kMoranWarps;   // 4 -> level-1 accumulation over 16 spots
constexpr int kSymRelaxIters = 8;
constexpr int kSymOffPad = 48;        // ints reserved per buffer for the chunk's record offsets (33 used)

__device__ __forceinline__ uint32_t sym_pair_hash(uint32_t i, uint32_t j) {
    const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
    uint32_t h = (lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

// own[i] bit a: spot i owns the pair {i, nbr[i][a]}.  Only first occurrences of a neighbour carry a
// bit; pairs only one endpoint lists are owned by that endpoint; quad-internal and self pairs carry
// no bit.  Mutual external pairs start with a hash-chosen owner.  cnt[i] = number of owned pairs.
__global__ void sym_init_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, uint64_t* __restrict__ own,
                                int32_t* __restrict__ cnt, int32_t* __restrict__ indeg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* A = nbr + i * k;
    uint64_t bits = 0;
    int c = 0;
    for (int a = 0; a < k; ++a) {
        const int32_t j = A[a];
        atomicAdd(indeg + j, 1);
        bool first = true;
        for (int b = 0; b < a; ++b) first &= (A[b] != j);
        if (!first) continue;
        if (j == (int32_t)i) { ++c; continue; }                        // self loop: one list entry, no bit
        if ((j >> 2) == (int32_t)(i >> 2)) continue;                   // quad-internal
        const int32_t* B = nbr + (int64_t)j * k;
        bool back = false;
        for (int b = 0; b < k; ++b) back |= (B[b] == (int32_t)i);
        bool mine = true;
        if (back) {
            const uint32_t lo = (uint32_t)(i < j ? i : j);
            mine = ((sym_pair_hash((uint32_t)i, (uint32_t)j) & 1u) ? lo : (uint32_t)(i ^ j ^ lo)) == (uint32_t)i;
        }
        if (mine) { ++c; bits |= 1ull << a; }
    }
    own[i] = bits;
    cnt[i] = c;
}

// one Jacobi sweep of owner balancing: a mutual pair moves to the other endpoint when that lowers
// the larger of the two list lengths; both endpoints evaluate the same rule on the same data.
__global__ void sym_relax_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int iter, const uint64_t* __restrict__ own_in,
                                 const int32_t* __restrict__ cnt_in, uint64_t* __restrict__ own_out, int32_t* __restrict__ cnt_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* A = nbr + i * k;
    const uint64_t mine_old = own_in[i];
    uint64_t bits = mine_old;
    int c = cnt_in[i];
    const int ci = c;
    for (int a = 0; a < k; ++a) {
        const int32_t j = A[a];
        if (j == (int32_t)i || (j >> 2) == (int32_t)(i >> 2)) continue;
        bool first = true;
        for (int b = 0; b < a; ++b) first &= (A[b] != j);
        if (!first) continue;
        const int32_t* B = nbr + (int64_t)j * k;
        int bpos = -1;
        for (int b = k - 1; b >= 0; --b) if (B[b] == (int32_t)i) bpos = b;    // first occurrence of i in N(j)
        if (bpos < 0) continue;                                                 // one-sided pair: fixed owner
        const bool i_owns = (mine_old >> a) & 1ull;
        const int cj = cnt_in[j];
        const int c_owner = i_owns ? ci : cj, c_other = i_owns ? cj : ci;
        const bool flip = (c_owner > c_other + 1) && ((sym_pair_hash((uint32_t)i, (uint32_t)j) >> (iter + 1)) & 1u);
        if (flip) {
            if (i_owns) { bits &= ~(1ull << a); --c; }
            else { bits |= 1ull << a; ++c; }
        }
    }
    own_out[i] = bits;
    cnt_out[i] = c;
}

// record size of every quad in 16-byte units
__global__ void sym_size_kernel(const int32_t* __restrict__ cnt, int64_t n, int64_t nquads, int32_t* __restrict__ qsize) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquads) return;
    int r = 0;
    for (int t = 0; t < 4; ++t) {
        const int64_t i = 4 * q + t;
        if (i < n && cnt[i] > r) r = cnt[i];
    }
    qsize[q] = (kSymHdrWords + kSymRoundWords * r) / 4;
}

// single-CTA exclusive scan of m ints; out[m] = total
__global__ void sym_scan_kernel(const int32_t* __restrict__ in, int64_t m, int32_t* __restrict__ out) {
    __shared__ int64_t sh[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int64_t per = (m + nt - 1) / nt;
    const int64_t b = (int64_t)t * per, e = b + per < m ? b + per : m;
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i) s += in[i];
    sh[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < nt; ++i) { int64_t v = sh[i]; sh[i] = run; run += v; }
        out[m] = (int32_t)run;
    }
    __syncthreads();
    int64_t run = sh[t];
    for (int64_t i = b; i < e; ++i) { out[i] = (int32_t)run; run += in[i]; }
}

// largest staged chunk (16-byte units): chunks start at multiples of kSymChunkQuads
__global__ void sym_chunkmax_kernel(const int32_t* __restrict__ off, int64_t nquads, int32_t* __restrict__ out_max) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t q0 = c * kSymChunkQuads;
    if (q0 >= nquads) return;
    const int64_t q1 = q0 + kSymChunkQuads < nquads ? q0 + kSymChunkQuads : nquads;
    atomicMax(out_max, off[q1] - off[q0]);
}

// one thread per quad writes its record
__global__ void sym_fill_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int64_t nquads, const uint64_t* __restrict__ own,
                                const int32_t* __restrict__ indeg, const int32_t* __restrict__ off, int32_t* __restrict__ rec_base) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquads) return;
    int32_t* rec = rec_base + (int64_t)off[q] * 4;
    const int R = ((off[q + 1] - off[q]) * 4 - kSymHdrWords) / kSymRoundWords;
    const int nvalid = (int)(n - 4 * q < 4 ? n - 4 * q : 4);
    float* recf = reinterpret_cast<float*>(rec);
    for (int w = 2; w < kSymHdrWords; ++w) recf[w] = 0.f;
    rec[0] = R;
    rec[1] = nvalid;
    for (int r = 0; r < R; ++r)
        for (int t = 0; t < 4; ++t) {
            const int64_t self = 4 * q + (t < nvalid ? t : 0);
            rec[kSymHdrWords + r * kSymRoundWords + t] = (int32_t)self;      // padding: a valid row, coefficient 0
            recf[kSymHdrWords + r * kSymRoundWords + 4 + t] = 0.f;
        }
    for (int t = 0; t < nvalid; ++t) {
        const int64_t i = 4 * q + t;
        const int32_t* A = nbr + i * k;
        const uint64_t bits = own[i];
        int slot = 0;
        float csum = 0.f;
        for (int a = 0; a < k; ++a) {
            const int32_t j = A[a];
            bool first = true;
            int mult = 0;
            for (int b = 0; b < k; ++b)
                if (A[b] == j) { if (b < a) first = false; ++mult; }
            if (!first) continue;
            float c;
            bool ext;
            if (j == (int32_t)i) {
                c = (float)mult; ext = true;                                   // a_ii z_i^2: own row with coefficient a_ii
            } else {
                const int32_t* B = nbr + (int64_t)j * k;
                int back = 0;
                for (int b = 0; b < k; ++b) back += (B[b] == (int32_t)i);
                c = (float)(mult + back);
                if ((j >> 2) == (int32_t)(i >> 2)) {
                    // quad-internal: written once, by the lower spot if it lists the pair, else by the upper one
                    const int u = t, v = (int)(j & 3);
                    const bool lower = u < v;
                    if (lower || back == 0) {
                        const int lo = lower ? u : v, hi = lower ? v : u;
                        const int idx = lo == 0 ? hi - 1 : (lo == 1 ? hi + 1 : 5);   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
                        recf[2 + idx] = c;
                    }
                    continue;
                }
                ext = (bits >> a) & 1ull;
            }
            if (!ext || slot >= R) continue;
            rec[kSymHdrWords + slot * kSymRoundWords + t] = j;
            recf[kSymHdrWords + slot * kSymRoundWords + 4 + t] = c;
            csum += c;
            ++slot;
        }
        recf[8 + t] = csum;
        recf[12 + t] = (float)indeg[i];
    }
}

__device__ __forceinline__ f2_t f2_dup(float c) { return f2_pack(c, c); }
__device__ __forceinline__ void r4_fmac(Row4& s, float c, const Row4& v) {   // s += c * v
    const f2_t cc = f2_dup(c);
    s.a = f2_fma(cc, v.a, s.a);
    s.b = f2_fma(cc, v.b, s.b);
}

// padded list slots have coefficient 0: their row load is skipped (predicated, no branch)
__device__ __forceinline__ void ld_row4_if(Row4& v, const char* p, float c) {
    asm("{\n"
        ".reg .pred q;\n"
        "setp.neu.f32 q, %3, 0f00000000;\n"
        "@q ld.global.nc.v2.b64 {%0, %1}, [%2];\n"
        "}\n"
        : "+l"(v.a), "+l"(v.b)
        : "l"(p), "f"(c));
}

template <bool GEARY, bool SHIFT, int MINB>
__global__ void __launch_bounds__(kMoranThreads, MINB)
moran_sym_kernel(const float* __restrict__ X, int64_t n, uint32_t row_bytes, int64_t ncol, const int32_t* __restrict__ plan_off,
                 const int4* __restrict__ plan_rec, int buf_units, const float* __restrict__ pilot, double* __restrict__ part,
                 int64_t gpad) {
    constexpr int NST = GEARY ? 5 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [2][kSymOffPad ints + buf_units x 16 B]  |  level-2 accumulators [NST][256] float4
    const int buf_bytes = kSymOffPad * 4 + buf_units * 16;
    float4* s_l2 = reinterpret_cast<float4*>(smem_raw + 2 * buf_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile_col = (int64_t)blockIdx.x * kGeneTile;
    int64_t col = tile_col + lane * 4;
    if (col >= ncol) col = tile_col;                                  // idle lanes shadow lane 0 (results unused)
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int64_t nquads = (n + 3) >> 2;
    const int64_t q_begin = (int64_t)blockIdx.y * (kV2SpotsPerCta / 4);
    const int64_t q_end = q_begin + kV2SpotsPerCta / 4 < nquads ? q_begin + kV2SpotsPerCta / 4 : nquads;
    const uint32_t tile_bytes = (uint32_t)((ncol - tile_col < kGeneTile ? ncol - tile_col : kGeneTile) * 4);

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
        s_l2[s * kMoranThreads + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // row registers of the two rounds in flight; a skipped (coefficient 0) load leaves the previous finite
    // row in place, which then contributes 0 * finite = 0
    Row4 va[4], vb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) va[t].a = va[t].b = vb[t].a = vb[t].b = z2;

    auto stage = [&](int buf, int64_t q0) {   // offsets and records of quads [q0, q0 + chunk) -> smem
        unsigned char* dst = smem_raw + buf * buf_bytes;
        const int nq = (int)(q_end - q0 < kSymChunkQuads ? q_end - q0 : kSymChunkQuads);
        if ((int)threadIdx.x <= nq) cp_async_4(dst + threadIdx.x * 4, plan_off + q0 + threadIdx.x);
        const int u0 = __ldg(plan_off + q0), u1 = __ldg(plan_off + q0 + nq);
        const int4* src = plan_rec + u0;
        int4* d4 = reinterpret_cast<int4*>(dst + kSymOffPad * 4);
        for (int u = threadIdx.x; u < u1 - u0; u += kMoranThreads) cp_async_16(d4 + u, src + u);
        cp_async_commit();
    };

    stage(0, q_begin);
    int buf = 0;
    for (int64_t q0 = q_begin; q0 < q_end; q0 += kSymChunkQuads, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();                                              // chunk `buf` visible; previous chunk fully consumed
        if (q0 + kSymChunkQuads < q_end) stage(buf ^ 1, q0 + kSymChunkQuads);
        const int32_t* s_off = reinterpret_cast<const int32_t*>(smem_raw + buf * buf_bytes);
        const int4* s_rec = reinterpret_cast<const int4*>(smem_raw + buf * buf_bytes + kSymOffPad * 4);
        const int off0 = s_off[0];
        const int64_t qw = q0 + warp * kSymQuadsPerWarp;
        // pull this warp's own rows of the next chunk towards L2 while the current one is consumed
        if (lane < 4 * kSymQuadsPerWarp) {
            const int64_t ip = 4 * (qw + kSymChunkQuads) + lane;
            if (ip < 4 * q_end && ip < n)
                l2_prefetch_bulk(reinterpret_cast<const char*>(X + tile_col) + (uint64_t)ip * row_bytes, tile_bytes);
        }
#pragma unroll 1
        for (int qi = 0; qi < kSymQuadsPerWarp; ++qi) {
            const int64_t q = qw + qi;
            if (q >= q_end) break;
            const int4* rec = s_rec + (s_off[warp * kSymQuadsPerWarp + qi] - off0);
            const int4 h0 = rec[0];
            const int R = h0.x, nvalid = h0.y;
            const uint32_t r0 = (uint32_t)(4 * q);
            Row4 x[4], s[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                x[t] = ld_row4(base + (uint64_t)(r0 + (t < nvalid ? t : 0)) * row_bytes);
                s[t].a = s[t].b = z2;
            }
            // two rounds per trip: 8 independent row loads in flight, then 16 multiply-adds
            int r = 0;
#pragma unroll 1
            for (; r + 1 < R; r += 2) {
                const int4 rows0 = rec[4 + 2 * r], rows1 = rec[6 + 2 * r];
                const float4 cf0 = *reinterpret_cast<const float4*>(rec + 5 + 2 * r);
                const float4 cf1 = *reinterpret_cast<const float4*>(rec + 7 + 2 * r);
                const int rr0[4] = {rows0.x, rows0.y, rows0.z, rows0.w}, rr1[4] = {rows1.x, rows1.y, rows1.z, rows1.w};
                const float cc0[4] = {cf0.x, cf0.y, cf0.z, cf0.w}, cc1[4] = {cf1.x, cf1.y, cf1.z, cf1.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) ld_row4_if(va[t], base + (uint64_t)(uint32_t)rr0[t] * row_bytes, cc0[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) ld_row4_if(vb[t], base + (uint64_t)(uint32_t)rr1[t] * row_bytes, cc1[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc0[t], va[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc1[t], vb[t]);
            }
            if (r < R) {
                const int4 rows0 = rec[4 + 2 * r];
                const float4 cf0 = *reinterpret_cast<const float4*>(rec + 5 + 2 * r);
                const int rr0[4] = {rows0.x, rows0.y, rows0.z, rows0.w};
                const float cc0[4] = {cf0.x, cf0.y, cf0.z, cf0.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) ld_row4_if(va[t], base + (uint64_t)(uint32_t)rr0[t] * row_bytes, cc0[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) r4_fmac(s[t], cc0[t], va[t]);
            }
            const float4 h1 = *reinterpret_cast<const float4*>(rec + 1);   // c03 c12 c13 c23
            const float4 cs = *reinterpret_cast<const float4*>(rec + 2);   // csum
            const float4 dg = *reinterpret_cast<const float4*>(rec + 3);   // indeg
            if (SHIFT) {
                const float ncs[4] = {-cs.x, -cs.y, -cs.z, -cs.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    r4_sub(x[t], p);
                    r4_fmac(s[t], ncs[t], p);
                }
            }
            // pairs inside the quad
            r4_fmac(s[0], __int_as_float(h0.z), x[1]);
            r4_fmac(s[0], __int_as_float(h0.w), x[2]);
            r4_fmac(s[0], h1.x, x[3]);
            r4_fmac(s[1], h1.y, x[2]);
            r4_fmac(s[1], h1.z, x[3]);
            r4_fmac(s[2], h1.w, x[3]);
            const float dgs[4] = {dg.x, dg.y, dg.z, dg.w};
            auto accumulate = [&](int t) {
                r4_add(l1[0], x[t]);
                r4_fma(l1[1], x[t], x[t]);
                r4_fma(l1[2], x[t], s[t]);
                if (GEARY) {
                    Row4 dx;
                    const f2_t dd = f2_dup(dgs[t]);
                    dx.a = f2_fma(dd, x[t].a, z2);
                    dx.b = f2_fma(dd, x[t].b, z2);
                    r4_add(l1[3], dx);
                    r4_fma(l1[4], dx, x[t]);
                } else {
                    r4_fmac(l1[3], dgs[t], x[t]);
                }
            };
            if (nvalid == 4) {                                        // every quad but the last one of a ragged n
#pragma unroll
                for (int t = 0; t < 4; ++t) accumulate(t);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t < nvalid) accumulate(t);
            }
        }
        // level-1 -> level-2 (every 16 spots of this warp)
#pragma unroll
        for (int st = 0; st < NST; ++st) {
            float4 acc = s_l2[st * kMoranThreads + threadIdx.x];
            float f0, f1, f2, f3;
            f2_unpack(l1[st].a, f0, f1);
            f2_unpack(l1[st].b, f2, f3);
            acc.x += f0; acc.y += f1; acc.z += f2; acc.w += f3;
            s_l2[st * kMoranThreads + threadIdx.x] = acc;
            l1[st].a = l1[st].b = z2;
        }
    }

    // level-2 -> fp64, cross-warp reduction in fixed order, one statistic at a time
    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_raw);               // [kMoranWarps][kGeneTile] (reuses the staging buffers)
#pragma unroll
    for (int st = 0; st < NST; ++st) {
        const float4 acc = s_l2[st * kMoranThreads + threadIdx.x];
        red[warp * kGeneTile + lane * 4 + 0] = (double)acc.x;
        red[warp * kGeneTile + lane * 4 + 1] = (double)acc.y;
        red[warp * kGeneTile + lane * 4 + 2] = (double)acc.z;
        red[warp * kGeneTile + lane * 4 + 3] = (double)acc.w;
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

// ---- finalize: fixed-order sum of the split partials, then the statistic in fp64 ------------
__global__ void moran_finalize_kernel(const double* __restrict__ part, int nsplit, int64_t gpad, int64_t g, int64_t n, int k,
                                      const float* __restrict__ pilot, bool geary, bool sym, double* __restrict__ out_I,
                                      double* __restrict__ out_C, double* __restrict__ out_stats) {
    int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= g) return;
    double S[kNStat] = {0, 0, 0, 0, 0};
    const int ns = geary ? kNStat : kNStat - 1;
    for (int sp = 0; sp < nsplit; ++sp)
        for (int st = 0; st < ns; ++st) S[st] += part[((int64_t)sp * kNStat + st) * gpad + col];
    const double dn = (double)n, dk = (double)k;
    const double m = S[0] / dn;                                   // mean of shifted values
    const double zz = S[1] - dn * m * m;                          // sum (x - mean)^2
    const double zwz = (S[2] - m * (dk * S[0] + S[3]) + m * m * dn * dk) / dk;  // z' W z, W = Adj / k
    // I = (n / s0) * z'Wz / z'z with s0 = n for row-standardised kNN weights
    out_I[col] = zwz / zz;
    // Geary numerator * k = sum_ij adj_ij (x_i - x_j)^2: accumulated directly (stream kernel) or from
    // k sum x^2 + sum_j indeg_j x_j^2 - 2 x'Ax (symmetric kernel; S[4] = sum_j indeg_j x_j^2)
    const double gnum = geary ? (sym ? dk * S[1] + S[4] - 2.0 * S[2] : S[4]) : 0.0;
    if (geary && out_C) out_C[col] = (dn - 1.0) * (gnum / dk) / (2.0 * dn * zz);
    if (out_stats) {
        out_stats[0 * g + col] = m + (double)pilot[col];   // pilot is all-zero for pre-shifted input
        out_stats[1 * g + col] = zz;
        out_stats[2 * g + col] = zwz;
        out_stats[3 * g + col] = gnum / dk;
    }
}

struct MoranPlan {
    int64_t gpad, nquads;
    int nsplit, ntile;
    size_t pilot_off, part_off, own_off[2], cnt_off[2], indeg_off, qsize_off, qoff_off, scal_off, rec_off, rec_bytes, total;
};

static MoranPlan moran_plan(int64_t n, int64_t ld, int k) {
    MoranPlan p;
    p.nsplit = (int)ceil_div<int64_t>(n, kV2SpotsPerCta);
    p.ntile = (int)ceil_div<int64_t>(ld, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    p.nquads = (n + 3) / 4;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 256); return r; };
    p.pilot_off = take((size_t)p.gpad * sizeof(float));
    p.part_off = take((size_t)p.nsplit * kNStat * p.gpad * sizeof(double));
    // symmetric-plan scratch: owner masks and list lengths (ping-pong), in-degrees, record sizes / offsets, records
    for (int b = 0; b < 2; ++b) { p.own_off[b] = take((size_t)n * 8); p.cnt_off[b] = take((size_t)n * 4); }
    p.indeg_off = take((size_t)n * 4);
    p.qsize_off = take((size_t)(p.nquads + 1) * 4);
    p.qoff_off = take((size_t)(p.nquads + 1) * 4);
    p.scal_off = take(256);
    // every pair is owned once and a spot owns at most k pairs: sum of rounds over quads <= n * k
    p.rec_bytes = (size_t)p.nquads * kSymHdrWords * 4 + (size_t)n * (size_t)(k < 64 ? k : 64) * kSymRoundWords * 4;
    p.rec_off = take(p.rec_bytes);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_moran_workspace_bytes(int64_t n_spots, int64_t n_genes, int k, unsigned flags) {
    (void)flags;
    if (n_spots <= 0 || n_genes <= 0) return 256;
    return chr::moran_plan(n_spots, chr::round_up<int64_t>(n_genes, 4), k > 0 ? k : 1).total + 512;
}

int chr_moran_dense_f32(const float* X, int64_t n_spots, int64_t n_genes, int64_t ld, const int32_t* nbr, int k,
                        unsigned flags, double* out_I, double* out_C, double* out_stats, void* workspace,
                        size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(n_spots > 1 && n_genes > 0, "chr_moran_dense_f32: need n_spots > 1 and n_genes > 0 (got %lld, %lld)",
                (long long)n_spots, (long long)n_genes);
    CHR_REQUIRE(k > 0 && k < n_spots, "chr_moran_dense_f32: need 0 < k < n_spots (k=%d)", k);
    CHR_REQUIRE(ld >= n_genes && ld % 4 == 0, "chr_moran_dense_f32: ld (%lld) must be >= n_genes and a multiple of 4", (long long)ld);
    CHR_REQUIRE(((uintptr_t)X & 15) == 0, "chr_moran_dense_f32: X must be 16-byte aligned");
    CHR_REQUIRE(X && nbr && out_I && workspace, "chr_moran_dense_f32: null pointer argument");
    CHR_REQUIRE(n_spots < (1LL << 31) && ld * 4 < (1LL << 32), "chr_moran_dense_f32: needs n_spots < 2^31, row bytes < 2^32");
    const bool geary = flags & CHR_FLAG_GEARY;
    CHR_REQUIRE(!geary || out_C, "chr_moran_dense_f32: CHR_FLAG_GEARY needs out_C");
    const bool timing = flags & CHR_FLAG_TIMING;

    // the kernels read round_up(n_genes, 4) <= ld columns
    const int64_t ncol = round_up<int64_t>(n_genes, 4);
    const unsigned variant = (flags >> CHR_VARIANT_SHIFT) & 0xffu;
    const bool preshifted = flags & CHR_FLAG_PRESHIFTED;
    // variant 0: symmetric quad kernel (v4, k <= 64).  variant 4 / k > 64: per-spot stream kernel (v2)
    const bool sym = variant != 4 && k <= 64;
    MoranPlan p = moran_plan(n_spots, ncol, k);
    CHR_REQUIRE(workspace_bytes >= p.total, "chr_moran_dense_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* pilot = (float*)(ws + p.pilot_off);
    double* part = (double*)(ws + p.part_off);

    if (preshifted) CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)p.gpad * sizeof(float), st));
    else moran_pilot_kernel<<<(unsigned)ceil_div<int64_t>(ncol, 256), 256, 0, st>>>(X, n_spots, ld, ncol, pilot);
    CHR_LAUNCH_CHECK();

    dim3 grid((unsigned)p.ntile, (unsigned)p.nsplit);
    const uint32_t rb = (uint32_t)(ld * 4);
    if (sym) {
        // ---- W -> symmetric quad plan (depends on W only)
        uint64_t* own[2] = {(uint64_t*)(ws + p.own_off[0]), (uint64_t*)(ws + p.own_off[1])};
        int32_t* cnt[2] = {(int32_t*)(ws + p.cnt_off[0]), (int32_t*)(ws + p.cnt_off[1])};
        int32_t* indeg = (int32_t*)(ws + p.indeg_off);
        int32_t* qsize = (int32_t*)(ws + p.qsize_off);
        int32_t* qoff = (int32_t*)(ws + p.qoff_off);
        int32_t* scal = (int32_t*)(ws + p.scal_off);
        int32_t* rec = (int32_t*)(ws + p.rec_off);
        const unsigned gs = (unsigned)ceil_div<int64_t>(n_spots, 128), gq = (unsigned)ceil_div<int64_t>(p.nquads, 128);
        CHR_CUDA(cudaMemsetAsync(indeg, 0, (size_t)n_spots * 4, st));
        CHR_CUDA(cudaMemsetAsync(scal, 0, 256, st));
        sym_init_kernel<<<gs, 128, 0, st>>>(nbr, n_spots, k, own[0], cnt[0], indeg);
        int cur = 0;
        for (int it = 0; it < kSymRelaxIters; ++it, cur ^= 1)
            sym_relax_kernel<<<gs, 128, 0, st>>>(nbr, n_spots, k, it, own[cur], cnt[cur], own[cur ^ 1], cnt[cur ^ 1]);
        sym_size_kernel<<<gq, 128, 0, st>>>(cnt[cur], n_spots, p.nquads, qsize);
        sym_scan_kernel<<<1, 1024, 0, st>>>(qsize, p.nquads, qoff);
        sym_chunkmax_kernel<<<(unsigned)ceil_div<int64_t>(ceil_div<int64_t>(p.nquads, kSymChunkQuads), 128), 128, 0, st>>>(qoff, p.nquads, scal);
        sym_fill_kernel<<<gq, 128, 0, st>>>(nbr, n_spots, k, p.nquads, own[cur], indeg, qoff, rec);
        CHR_LAUNCH_CHECK();
        int chunk_units = 0;
        CHR_CUDA(cudaMemcpyAsync(&chunk_units, scal, sizeof(int), cudaMemcpyDeviceToHost, st));
        CHR_CUDA(cudaStreamSynchronize(st));
        const int buf_units = chunk_units < 256 ? 256 : chunk_units;      // the final reduction reuses >= 8 KB of the buffers
        const int nst = geary ? 5 : 4;
        const size_t smem = (size_t)2 * (kSymOffPad * 4 + (size_t)buf_units * 16) + (size_t)nst * kMoranThreads * 16;
        CHR_REQUIRE(smem <= 200 * 1024, "chr_moran_dense_f32: neighbour lists too long for the staged plan (%zu B of shared memory)", smem);
        timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_V4(GG, SH)                                                                                              \
    do {                                                                                                            \
        if (variant == 5) CHR_V4M(GG, SH, 1); else CHR_V4M(GG, SH, 2);                                             \
    } while (0)
#define CHR_V4M(GG, SH, MB)                                                                                         \
    do {                                                                                                            \
        auto kern = moran_sym_kernel<GG, SH, MB>;                                                                    \
        CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        kern<<<grid, kMoranThreads, smem, st>>>(X, n_spots, rb, ncol, qoff, (const int4*)rec, buf_units, pilot, part, p.gpad); \
    } while (0)
        if (geary && preshifted) CHR_V4(true, false);
        else if (geary) CHR_V4(true, true);
        else if (preshifted) CHR_V4(false, false);
        else CHR_V4(false, true);
#undef CHR_V4
#undef CHR_V4M
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

    moran_finalize_kernel<<<(unsigned)ceil_div<int64_t>(n_genes, 256), 256, 0, st>>>(
        part, p.nsplit, p.gpad, n_genes, n_spots, k, pilot, geary, sym, out_I, out_C, out_stats);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
