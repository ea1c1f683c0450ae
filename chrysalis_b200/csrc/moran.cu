// K3/K4 - Moran's I (and Geary's C) for every gene of a dense spots x genes matrix, one pass.
//
// Replaces the per-gene loop  esda.moran.Moran(gene_matrix[c], w, permutations=0).I
// [+ esda.geary.Geary(...).C]  of /root/reference/chrysalis/core.py:71-76.
//
// Layout: X is [n_spots][ld] float32 row-major (row = spot, spots in a spatially coherent
// order), W is an ELL kNN list nbr[n_spots][k] with implicit weights 1/k (libpysal
// transform 'R').  A warp owns one spot at a time and 128 consecutive genes of it (one
// float4 per lane): the own row and the k neighbour rows are coalesced 512-byte segments,
// every lane accumulates its four genes over the spots of the CTA's spot range, so there
// is no cross-thread reduction in the streaming loop.  Re-reads of neighbour rows are
// served by L1/L2 (spots are Hilbert-ordered by the host), HBM sees X once.
//
// Numerics: values are shifted by a per-gene pilot mean p (mean of a fixed
// spot sample) before any product, so the one-pass identities
//     z'z  = S2 - n m'^2,      k z'Wz = SW - m' (k S1 + SS) + m'^2 n k       (m' = S1/n)
// with S1 = sum x', S2 = sum x'^2, SW = sum_i x'_i s_i, SS = sum_i s_i, s_i = sum_{j in N(i)} x'_j
// carry no cancellation; partial sums are fp32 over 16 spots per warp, then fp64.
#include "common.cuh"

namespace chr {

constexpr int kMoranThreads = 256;
constexpr int kMoranWarps = kMoranThreads / 32;
constexpr int kGeneTile = 128;   // genes per CTA = 32 lanes x float4
constexpr int kChunk = 16;       // spots per warp between fp32 -> fp64 flushes
constexpr int kNStat = 5;        // S1, S2, SW, SS, SQ (Geary numerator * k)
constexpr int kPilotSample = 512;

__host__ __device__ inline int64_t moran_spots_per_split(int64_t n) {
    // depends on n only: the summation tree of a gene is identical however genes are sharded
    int64_t s = round_up<int64_t>(ceil_div<int64_t>(n, 64), 64);
    return s < 512 ? 512 : s;
}

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

__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void f4_fma(float4& a, const float4& b, const float4& c) {
    a.x = fmaf(b.x, c.x, a.x); a.y = fmaf(b.y, c.y, a.y); a.z = fmaf(b.z, c.z, a.z); a.w = fmaf(b.w, c.w, a.w);
}
__device__ __forceinline__ void f4_sqdiff(float4& a, const float4& x, const float4& v) {
    float d;
    d = x.x - v.x; a.x = fmaf(d, d, a.x);
    d = x.y - v.y; a.y = fmaf(d, d, a.y);
    d = x.z - v.z; a.z = fmaf(d, d, a.z);
    d = x.w - v.w; a.w = fmaf(d, d, a.w);
}

struct Acc4 {
    double v[4];
    __device__ __forceinline__ void flush(float4& f) {
        v[0] += (double)f.x; v[1] += (double)f.y; v[2] += (double)f.z; v[3] += (double)f.w;
        f = make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

// one elected lane asks the TMA unit to pull a row segment into L2 ahead of use
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// K = compile-time neighbour count (0: runtime k, any value); MINB = CTAs/SM the register
// budget is sized for; PF = L2 prefetch distance in warp iterations (0 = off)
template <int K, bool GEARY, int MINB, int PF>
__global__ void __launch_bounds__(kMoranThreads, MINB)
moran_ell_kernel(const float* __restrict__ X, int64_t n, int64_t ld, int64_t ncol, const int32_t* __restrict__ nbr,
                 int k_rt, const float* __restrict__ pilot, int64_t spots_per_split, double* __restrict__ part,
                 int64_t gpad) {
    const int k = K ? K : k_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t col = (int64_t)blockIdx.x * kGeneTile + lane * 4;
    const bool active = col < ncol;   // ncol % 4 == 0: a lane is in or out as a whole
    const int64_t s_begin = (int64_t)blockIdx.y * spots_per_split;
    const int64_t s_end = s_begin + spots_per_split < n ? s_begin + spots_per_split : n;
    const float* __restrict__ Xc = X + (active ? col : 0);

    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) p = ldg_f4(pilot + col);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 f1 = zero, f2 = zero, fw = zero, fs = zero, fq = zero;
    Acc4 a1{}, a2{}, aw{}, as{}, aq{};
    int cnt = 0;

    // bytes of this CTA's 128-gene slab that exist in the row (16-byte multiple)
    const int64_t tile_col = (int64_t)blockIdx.x * kGeneTile;
    const uint32_t tile_bytes = (uint32_t)((ncol - tile_col < kGeneTile ? ncol - tile_col : kGeneTile) * 4);

    for (int64_t i = s_begin + warp; i < s_end; i += kMoranWarps) {
        if (PF > 0 && lane == 0) {
            const int64_t ip = i + (int64_t)PF * kMoranWarps;
            if (ip < s_end) l2_prefetch_bulk(X + ip * ld + tile_col, tile_bytes);
        }
        float4 x = zero, s = zero;
        if (active) x = ldg_f4(Xc + i * ld);
        if constexpr (K > 0 && K <= 32) {
            int my = 0;
            if (lane < K) my = __ldg(nbr + i * K + lane);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int nj = __shfl_sync(0xffffffffu, my, j);
                if (active) {
                    float4 v = ldg_f4(Xc + (int64_t)nj * ld);
                    if (GEARY) f4_sqdiff(fq, x, v);
                    v.x -= p.x; v.y -= p.y; v.z -= p.z; v.w -= p.w;
                    f4_add(s, v);
                }
            }
        } else {
            for (int jb = 0; jb < k; jb += 32) {
                int my = 0;
                if (jb + lane < k) my = __ldg(nbr + i * k + jb + lane);
                const int kk = k - jb < 32 ? k - jb : 32;
#pragma unroll 4
                for (int j = 0; j < kk; ++j) {
                    const int nj = __shfl_sync(0xffffffffu, my, j);
                    if (active) {
                        float4 v = ldg_f4(Xc + (int64_t)nj * ld);
                        if (GEARY) f4_sqdiff(fq, x, v);
                        v.x -= p.x; v.y -= p.y; v.z -= p.z; v.w -= p.w;
                        f4_add(s, v);
                    }
                }
            }
        }
        x.x -= p.x; x.y -= p.y; x.z -= p.z; x.w -= p.w;
        f4_add(f1, x);
        f4_fma(f2, x, x);
        f4_fma(fw, x, s);
        f4_add(fs, s);
        if (++cnt == kChunk) {
            cnt = 0;
            a1.flush(f1); a2.flush(f2); aw.flush(fw); as.flush(fs);
            if (GEARY) aq.flush(fq);
        }
    }
    a1.flush(f1); a2.flush(f2); aw.flush(fw); as.flush(fs);
    if (GEARY) aq.flush(fq);

    // cross-warp reduction in fixed order, one statistic at a time
    __shared__ double red[kMoranWarps][kGeneTile];
    auto reduce_store = [&](const Acc4& a, int st) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[warp][lane * 4 + c] = a.v[c];
        __syncthreads();
        if (threadIdx.x < kGeneTile) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kMoranWarps; ++w) t += red[w][threadIdx.x];
            const int64_t gcol = (int64_t)blockIdx.x * kGeneTile + threadIdx.x;
            part[((int64_t)blockIdx.y * kNStat + st) * gpad + gcol] = t;
        }
        __syncthreads();
    };
    reduce_store(a1, 0);
    reduce_store(a2, 1);
    reduce_store(aw, 2);
    reduce_store(as, 3);
    if (GEARY) reduce_store(aq, 4);
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
// v3 pair kernel.  The r1b capture showed the v2 kernel pinned on the L1 data pipe (72 % of its
// wavefront peak: 9 x 512-byte row reads per spot).  Consecutive spots of the Hilbert order
// are spatial neighbours and share most of their neighbourhood (union of two closed
// neighbourhoods: 12.1 rows instead of 18 on a square lattice, 10 on a hex lattice, 12.1 on
// random points), so spots are processed in PAIRS (a, b) = (2q, 2q+1) and W is re-expressed
// per pair as three disjoint lists
//        only-a | common | only-b        (+ flags  b in N(a),  a in N(b))
//   s_a = sum(only-a) + sum(common) [+ x_b],   s_b = sum(only-b) + sum(common) [+ x_a]
// which needs 6 row reads and 5.5 adds per spot instead of 9 and 8.  The plan is built from
// the kNN lists by moran_pair_plan_kernel; lists with duplicate entries or odd shapes take a
// generic per-spot path (results are identical either way, only the summation order differs).
// =============================================================================================
constexpr int kPairRecPad = 4;   // ints before the index list: meta, 3 x pad (keeps lists 16-byte aligned)
__host__ __device__ constexpr int pair_stride(int k) { return (2 * k + kPairRecPad + 3) & ~3; }   // ints per pair record

__host__ __device__ constexpr int pair_ncm(int K) { return K / 2; }
__host__ __device__ constexpr int pair_nam(int K) { return K / 2 + (K <= 4 ? 1 : 0); }

// meta: n_a | n_c << 8 | n_b << 16 | fa << 24 | fb << 25 | fast << 26 | has_b << 27
__global__ void moran_pair_plan_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int ncm, int nam,
                                       int32_t* __restrict__ plan) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npairs = (n + 1) / 2;
    if (q >= npairs) return;
    const int64_t a = 2 * q, b = 2 * q + 1;
    const bool hasb = b < n;
    const int stride = pair_stride(k);
    int32_t* rec = plan + q * stride;
    const int32_t* A = nbr + a * k;
    const int32_t* B = nbr + (hasb ? b : a) * k;
    bool simple = hasb;
    for (int i = 0; i < k && simple; ++i)
        for (int j = i + 1; j < k; ++j)
            if (A[i] == A[j] || B[i] == B[j]) { simple = false; break; }
    int na = 0, nc = 0, nb = 0, fa = 0, fb = 0;
    int32_t* out = rec + kPairRecPad;
    if (!hasb) {
        for (int i = 0; i < k; ++i) out[na++] = A[i];
    } else if (!simple) {
        for (int i = 0; i < k; ++i) out[na++] = A[i];
        for (int i = 0; i < k; ++i) out[na + nb++] = B[i];
    } else {
        // only-a
        for (int i = 0; i < k; ++i) {
            const int32_t x = A[i];
            if (x == (int32_t)b) { fa = 1; continue; }
            bool inB = false;
            for (int j = 0; j < k; ++j) inB |= (B[j] == x);
            if (!inB) out[na++] = x;
        }
        // common (order of A)
        for (int i = 0; i < k; ++i) {
            const int32_t x = A[i];
            if (x == (int32_t)b) continue;
            bool inB = false;
            for (int j = 0; j < k; ++j) inB |= (B[j] == x);
            if (inB) out[na + nc++] = x;
        }
        // only-b
        for (int i = 0; i < k; ++i) {
            const int32_t x = B[i];
            if (x == (int32_t)a) { fb = 1; continue; }
            bool inA = false;
            for (int j = 0; j < k; ++j) inA |= (A[j] == x && x != (int32_t)b);
            if (!inA) out[na + nc + nb++] = x;
        }
    }
    for (int i = na + nc + nb; i < stride - kPairRecPad; ++i) out[i] = (int32_t)a;   // padding: a valid row
    const int fast = (hasb && simple && nc <= ncm && na <= nam && nb <= nam) ? 1 : 0;
    rec[0] = na | (nc << 8) | (nb << 16) | (fa << 24) | (fb << 25) | (fast << 26) | ((hasb ? 1 : 0) << 27);
    rec[1] = rec[2] = rec[3] = 0;
}

constexpr int kV3PairsPerWarp = 8;                               // 16 spots: level-1 accumulation length
constexpr int kV3ChunkPairs = kV3PairsPerWarp * kMoranWarps;     // 64 pairs = 128 spots per CTA chunk

template <int K, bool SHIFT, int MINB>
__global__ void __launch_bounds__(kMoranThreads, MINB)
moran_pair_kernel(const float* __restrict__ X, int64_t n, uint32_t row_bytes, int64_t ncol, const int32_t* __restrict__ plan,
                  int k_rt, const float* __restrict__ pilot, double* __restrict__ part, int64_t gpad) {
    const int k = K ? K : k_rt;
    const int stride = pair_stride(k);                            // ints per pair record (multiple of 4)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* s_plan = reinterpret_cast<int32_t*>(smem_raw);       // [2][kV3ChunkPairs * stride]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile_col = (int64_t)blockIdx.x * kGeneTile;
    int64_t col = tile_col + lane * 4;
    if (col >= ncol) col = tile_col;
    const char* __restrict__ base = reinterpret_cast<const char*>(X + col);
    const int64_t s_begin = (int64_t)blockIdx.y * kV2SpotsPerCta;
    const int64_t s_end = s_begin + kV2SpotsPerCta < n ? s_begin + kV2SpotsPerCta : n;
    const int64_t q_begin = s_begin / 2, q_end = (s_end + 1) / 2;  // kV2SpotsPerCta is even
    const uint32_t tile_bytes = (uint32_t)((ncol - tile_col < kGeneTile ? ncol - tile_col : kGeneTile) * 4);

    Row4 p;
    p.a = p.b = f2_pack(0.f, 0.f);
    if (SHIFT) {
        const float4 pf = __ldg(reinterpret_cast<const float4*>(pilot + col));
        p.a = f2_pack(pf.x, pf.y);
        p.b = f2_pack(pf.z, pf.w);
    }
    const f2_t z2 = f2_pack(0.f, 0.f);
    Row4 l1[4], l2[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { l1[s].a = l1[s].b = z2; l2[s].a = l2[s].b = z2; }

    auto ld_shifted = [&](uint32_t row) {
        Row4 v = ld_row4(base + (uint64_t)row * row_bytes);
        if (SHIFT) r4_sub(v, p);
        return v;
    };
    auto accumulate = [&](const Row4& x, const Row4& s) {
        r4_add(l1[0], x);
        r4_fma(l1[1], x, x);
        r4_fma(l1[2], x, s);
        r4_add(l1[3], s);
    };
    // each warp stages the records of its own 8 pairs of a chunk (no CTA barrier on the way)
    const int warp_ints = kV3PairsPerWarp * stride;
    auto stage = [&](int buf, int64_t q0) {
        int32_t* dst = s_plan + (buf * kV3ChunkPairs + warp * kV3PairsPerWarp) * stride;
        const int64_t qw = q0 + warp * kV3PairsPerWarp;
        int64_t rem = (q_end - qw) * stride;
        if (rem > warp_ints) rem = warp_ints;
        const int32_t* src = plan + qw * stride;
        for (int t = lane * 4; t < rem; t += 128) cp_async_16(dst + t, src + t);
        cp_async_commit();
    };

    stage(0, q_begin);
    int buf = 0, chunk_no = 0;
    for (int64_t q0 = q_begin; q0 < q_end; q0 += kV3ChunkPairs, buf ^= 1, ++chunk_no) {
        cp_async_wait_all();
        __syncwarp();
        if ((chunk_no & 3) == 3) __syncthreads();                 // keep the warps of a CTA within 4 chunks of each other (L1 sharing)
        if (q0 + kV3ChunkPairs < q_end) stage(buf ^ 1, q0 + kV3ChunkPairs);
        const int32_t* my = s_plan + (buf * kV3ChunkPairs + warp * kV3PairsPerWarp) * stride;
        const int64_t qw = q0 + warp * kV3PairsPerWarp;
        if (lane < 2 * kV3PairsPerWarp) {                          // own rows of this warp's slice of the next chunk -> L2
            const int64_t ip = 2 * (qw + kV3ChunkPairs) + lane;
            if (ip < s_end) l2_prefetch_bulk(reinterpret_cast<const char*>(X + tile_col) + (uint64_t)ip * row_bytes, tile_bytes);
        }
#pragma unroll 1
        for (int pi = 0; pi < kV3PairsPerWarp; ++pi) {
            const int64_t q = qw + pi;
            if (q >= q_end) break;
            const int32_t* rec = my + pi * stride;
            const int meta = rec[0];
            const int na = meta & 0xff, nc = (meta >> 8) & 0xff, nb = (meta >> 16) & 0xff;
            const bool fa = (meta >> 24) & 1, fb = (meta >> 25) & 1, fast = (meta >> 26) & 1, hasb = (meta >> 27) & 1;
            const int32_t* idx = rec + kPairRecPad;
            const uint32_t ra = (uint32_t)(2 * q), rb = hasb ? ra + 1 : ra;
            const Row4 xa = ld_shifted(ra);
            const Row4 xb = ld_shifted(rb);
            Row4 sa, sb;
            sa.a = sa.b = sb.a = sb.b = z2;
            bool done = false;
            if constexpr (K > 0) {
                if (fast) {
                    constexpr int NCM = pair_ncm(K), NAM = pair_nam(K);
                    Row4 va[NAM], vc[NCM], vb[NAM];
#pragma unroll
                    for (int j = 0; j < NAM; ++j) if (j < na) va[j] = ld_shifted((uint32_t)idx[j]);
#pragma unroll
                    for (int j = 0; j < NCM; ++j) if (j < nc) vc[j] = ld_shifted((uint32_t)idx[na + j]);
#pragma unroll
                    for (int j = 0; j < NAM; ++j) if (j < nb) vb[j] = ld_shifted((uint32_t)idx[na + nc + j]);
                    Row4 sc;
                    sc.a = sc.b = z2;
#pragma unroll
                    for (int j = 0; j < NCM; ++j) if (j < nc) r4_add(sc, vc[j]);
                    sa = sc;
                    sb = sc;
#pragma unroll
                    for (int j = 0; j < NAM; ++j) if (j < na) r4_add(sa, va[j]);
#pragma unroll
                    for (int j = 0; j < NAM; ++j) if (j < nb) r4_add(sb, vb[j]);
                    if (fa) r4_add(sa, xb);
                    if (fb) r4_add(sb, xa);
                    done = true;
                }
            }
            if (!done) {   // generic per-spot path: a reads only-a|common, b reads common|only-b
                const int la = na + nc, lb = nc + nb;
#pragma unroll 4
                for (int j = 0; j < la; ++j) r4_add(sa, ld_shifted((uint32_t)idx[j]));
#pragma unroll 4
                for (int j = 0; j < lb; ++j) r4_add(sb, ld_shifted((uint32_t)idx[na + j]));
                if (fa) r4_add(sa, xb);
                if (fb) r4_add(sb, xa);
            }
            accumulate(xa, sa);
            if (hasb) accumulate(xb, sb);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            r4_add(l2[s], l1[s]);
            l1[s].a = l1[s].b = z2;
        }
    }

    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_raw);
#pragma unroll
    for (int st = 0; st < 4; ++st) {
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

// ---- finalize: fixed-order sum of the split partials, then the statistic in fp64 ------------
__global__ void moran_finalize_kernel(const double* __restrict__ part, int nsplit, int64_t gpad, int64_t g, int64_t n, int k,
                                      const float* __restrict__ pilot, bool geary, double* __restrict__ out_I,
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
    if (geary && out_C) out_C[col] = (dn - 1.0) * (S[4] / dk) / (2.0 * dn * zz);
    if (out_stats) {
        out_stats[0 * g + col] = m + (double)pilot[col];   // pilot is all-zero for pre-shifted input
        out_stats[1 * g + col] = zz;
        out_stats[2 * g + col] = zwz;
        out_stats[3 * g + col] = geary ? S[4] / dk : 0.0;
    }
}

struct MoranPlan {
    int64_t spots_per_split, gpad;
    int nsplit, ntile;
    size_t pilot_off, part_off, pair_off, total;
};

static MoranPlan moran_plan(int64_t n, int64_t g, int64_t ld, bool v2, int k) {
    MoranPlan p;
    p.spots_per_split = v2 ? kV2SpotsPerCta : moran_spots_per_split(n);
    p.nsplit = (int)ceil_div<int64_t>(n, p.spots_per_split);
    p.ntile = (int)ceil_div<int64_t>(ld, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    p.pilot_off = 0;
    p.part_off = round_up<size_t>((size_t)p.gpad * sizeof(float), 256);
    p.pair_off = round_up<size_t>(p.part_off + (size_t)p.nsplit * kNStat * p.gpad * sizeof(double), 256);
    p.total = p.pair_off + (size_t)((n + 1) / 2) * pair_stride(k) * sizeof(int32_t);
    return p;
}

template <bool GEARY, int MINB, int PF>
static void launch_ell(int k, dim3 grid, cudaStream_t st, const float* X, int64_t n, int64_t ld, int64_t ncol,
                       const int32_t* nbr, const float* pilot, int64_t sps, double* part, int64_t gpad) {
#define CHR_ELL(KK) moran_ell_kernel<KK, GEARY, MINB, PF><<<grid, kMoranThreads, 0, st>>>(X, n, ld, ncol, nbr, k, pilot, sps, part, gpad)
    switch (k) {
        case 4: CHR_ELL(4); break;
        case 6: CHR_ELL(6); break;
        case 8: CHR_ELL(8); break;
        default: CHR_ELL(0); break;
    }
#undef CHR_ELL
}

}  // namespace chr

extern "C" {

size_t chr_moran_workspace_bytes(int64_t n_spots, int64_t n_genes, int k, unsigned flags) {
    (void)flags;
    if (n_spots <= 0 || n_genes <= 0) return 256;
    const int64_t ldc = chr::round_up<int64_t>(n_genes, 4);
    const int kk = k > 0 ? k : 1;
    const size_t a = chr::moran_plan(n_spots, n_genes, ldc, false, kk).total, b = chr::moran_plan(n_spots, n_genes, ldc, true, kk).total;
    return (a > b ? a : b) + 512;
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
    const bool geary = flags & CHR_FLAG_GEARY;
    CHR_REQUIRE(!geary || out_C, "chr_moran_dense_f32: CHR_FLAG_GEARY needs out_C");
    const bool timing = flags & CHR_FLAG_TIMING;

    // the kernel reads ld columns; statistics are kept for the first round_up(n_genes,4) <= ld
    const int64_t ldc = round_up<int64_t>(n_genes, 4);
    const unsigned variant = (flags >> CHR_VARIANT_SHIFT) & 0xffu;
    const bool preshifted = flags & CHR_FLAG_PRESHIFTED;
    // variant 0/4: stream kernel (v2, default).  1: v1 ELL kernel, 6/7: v3 pair kernel (kept for regression / tuning)
    const bool v1 = variant == 1 && !geary;
    const bool v3 = !geary && (variant == 6 || variant == 7);
    const bool v2 = !v1;   // v2 and v3 share the 2048-spot split structure
    CHR_REQUIRE(!preshifted || !v1, "chr_moran_dense_f32: CHR_FLAG_PRESHIFTED is not supported by the v1 kernel");
    MoranPlan p = moran_plan(n_spots, n_genes, ldc, v2, k);
    CHR_REQUIRE(workspace_bytes >= p.total, "chr_moran_dense_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* pilot = (float*)(ws + p.pilot_off);
    double* part = (double*)(ws + p.part_off);

    // columns the kernels touch: [0, ncol), ncol % 4 == 0, ncol <= ld and <= gpad
    const int64_t ncol = ldc;
    if (preshifted) CHR_CUDA(cudaMemsetAsync(pilot, 0, (size_t)p.gpad * sizeof(float), st));
    else moran_pilot_kernel<<<(unsigned)ceil_div<int64_t>(ncol, 256), 256, 0, st>>>(X, n_spots, ld, ncol, pilot);
    CHR_LAUNCH_CHECK();

    dim3 grid((unsigned)p.ntile, (unsigned)p.nsplit);
#define CHR_ARGS k, grid, st, X, n_spots, ld, ncol, nbr, pilot, p.spots_per_split, part, p.gpad
    if (v3) {
        CHR_REQUIRE(n_spots < (1LL << 31) && ld * 4 < (1LL << 32), "chr_moran_dense_f32: pair kernel needs n_spots < 2^31, row bytes < 2^32");
        CHR_REQUIRE(k <= 120, "chr_moran_dense_f32: pair kernel supports k <= 120");
        int32_t* pair_plan = (int32_t*)(ws + p.pair_off);
        const int64_t npairs = (n_spots + 1) / 2;
        const int ncm = (k == 4 || k == 6 || k == 8) ? pair_ncm(k) : -1, nam = (k == 4 || k == 6 || k == 8) ? pair_nam(k) : -1;
        moran_pair_plan_kernel<<<(unsigned)ceil_div<int64_t>(npairs, 128), 128, 0, st>>>(nbr, n_spots, k, ncm, nam, pair_plan);
        CHR_LAUNCH_CHECK();
        size_t smem = (size_t)2 * kV3ChunkPairs * pair_stride(k) * sizeof(int32_t);
        if (smem < sizeof(double) * kMoranWarps * kGeneTile) smem = sizeof(double) * kMoranWarps * kGeneTile;
        const uint32_t rb = (uint32_t)(ld * 4);
        timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_V3(KK, SH, MB)                                                                                          \
    do {                                                                                                            \
        auto kern = moran_pair_kernel<KK, SH, MB>;                                                                 \
        if (smem > 48 * 1024) CHR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kMoranThreads, smem, st>>>(X, n_spots, rb, ncol, pair_plan, k, pilot, part, p.gpad);         \
    } while (0)
#define CHR_V3K(SH, MB)                                  \
    switch (k) {                                         \
        case 4: CHR_V3(4, SH, MB); break;                \
        case 6: CHR_V3(6, SH, MB); break;                \
        case 8: CHR_V3(8, SH, MB); break;                \
        default: CHR_V3(0, SH, MB); break;               \
    }
        if (preshifted) { CHR_V3K(false, 2) }
        else if (variant == 7) { CHR_V3K(true, 1) }
        else { CHR_V3K(true, 2) }
#undef CHR_V3K
#undef CHR_V3
    } else if (v2) {
        CHR_REQUIRE(n_spots < (1LL << 32) && ld * 4 < (1LL << 32), "chr_moran_dense_f32: v2 kernel needs n_spots, row bytes < 2^32");
        const int vec16 = (k % 4 == 0) && (((uintptr_t)nbr & 15) == 0);
        size_t smem = (size_t)2 * kV2Chunk * k * sizeof(int32_t);
        if (smem < sizeof(double) * kMoranWarps * kGeneTile) smem = sizeof(double) * kMoranWarps * kGeneTile;
        const uint32_t rb = (uint32_t)(ld * 4);
        timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_V2(KK, GG, MB, SH)                                                                                      \
    do {                                                                                                            \
        auto kern = moran_stream_kernel<KK, GG, MB, SH>;                                                               \
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
    } else {
        timer_begin(CHR_FAMILY_MORAN, timing, st);
        launch_ell<false, 2, 8>(CHR_ARGS);
    }
#undef CHR_ARGS
    timer_end(CHR_FAMILY_MORAN, timing, st);
    CHR_LAUNCH_CHECK();

    moran_finalize_kernel<<<(unsigned)ceil_div<int64_t>(n_genes, 256), 256, 0, st>>>(
        part, p.nsplit, p.gpad, n_genes, n_spots, k, pilot, geary, out_I, out_C, out_stats);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
