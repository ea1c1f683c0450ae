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
    const float kf = (float)k;
    const float4 kp = make_float4(kf * p.x, kf * p.y, kf * p.z, kf * p.w);

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
                    const float4 v = ldg_f4(Xc + (int64_t)nj * ld);
                    f4_add(s, v);
                    if (GEARY) f4_sqdiff(fq, x, v);
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
                        const float4 v = ldg_f4(Xc + (int64_t)nj * ld);
                        f4_add(s, v);
                        if (GEARY) f4_sqdiff(fq, x, v);
                    }
                }
            }
        }
        x.x -= p.x; x.y -= p.y; x.z -= p.z; x.w -= p.w;
        s.x -= kp.x; s.y -= kp.y; s.z -= kp.z; s.w -= kp.w;
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
        out_stats[0 * g + col] = m + (double)pilot[col];
        out_stats[1 * g + col] = zz;
        out_stats[2 * g + col] = zwz;
        out_stats[3 * g + col] = geary ? S[4] / dk : 0.0;
    }
}

struct MoranPlan {
    int64_t spots_per_split, gpad;
    int nsplit, ntile;
    size_t pilot_off, part_off, total;
};

static MoranPlan moran_plan(int64_t n, int64_t g, int64_t ld) {
    MoranPlan p;
    p.spots_per_split = moran_spots_per_split(n);
    p.nsplit = (int)ceil_div<int64_t>(n, p.spots_per_split);
    p.ntile = (int)ceil_div<int64_t>(ld, kGeneTile);
    p.gpad = (int64_t)p.ntile * kGeneTile;
    p.pilot_off = 0;
    p.part_off = round_up<size_t>((size_t)p.gpad * sizeof(float), 256);
    p.total = p.part_off + (size_t)p.nsplit * kNStat * p.gpad * sizeof(double);
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
    (void)k; (void)flags;
    if (n_spots <= 0 || n_genes <= 0) return 256;
    return chr::moran_plan(n_spots, n_genes, chr::round_up<int64_t>(n_genes, 4)).total + 512;
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
    MoranPlan p = moran_plan(n_spots, n_genes, ldc);
    CHR_REQUIRE(workspace_bytes >= p.total, "chr_moran_dense_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* pilot = (float*)(ws + p.pilot_off);
    double* part = (double*)(ws + p.part_off);

    // columns the kernels touch: [0, ncol), ncol % 4 == 0, ncol <= ld and <= gpad
    const int64_t ncol = ldc;
    moran_pilot_kernel<<<(unsigned)ceil_div<int64_t>(ncol, 256), 256, 0, st>>>(X, n_spots, ld, ncol, pilot);
    CHR_LAUNCH_CHECK();

    dim3 grid((unsigned)p.ntile, (unsigned)p.nsplit);
    timer_begin(CHR_FAMILY_MORAN, timing, st);
#define CHR_ARGS k, grid, st, X, n_spots, ld, ncol, nbr, pilot, p.spots_per_split, part, p.gpad
    const unsigned variant = (flags >> CHR_VARIANT_SHIFT) & 0xffu;
    if (geary) launch_ell<true, 2, 8>(CHR_ARGS);
    else switch (variant) {
        case 1: launch_ell<false, 2, 0>(CHR_ARGS); break;   // no prefetch
        case 2: launch_ell<false, 2, 32>(CHR_ARGS); break;  // far prefetch
        case 3: launch_ell<false, 3, 8>(CHR_ARGS); break;   // 3 CTAs/SM
        default: launch_ell<false, 2, 8>(CHR_ARGS); break;
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
