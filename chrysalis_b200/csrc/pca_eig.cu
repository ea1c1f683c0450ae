// K7 / K8 - top-r eigenpairs of the SVG covariance and the PCA scores.
//
// Replaces the ARPACK part of  PCA(n_components, svd_solver='arpack', random_state=42)
// (/root/reference/chrysalis/core.py:127-128; sklearn _pca.py:_fit_truncated -> scipy svds).
// The covariance C (s x s, fp64, s <= a few thousand) comes from pca_gram.cu.
//
// Solver: Lanczos with full (CGS2) reorthogonalisation in fp64 on C, m steps; the symmetric
// tridiagonal T_m is solved for its r largest eigenvalues by Sturm bisection and for the
// eigenvectors by inverse iteration (partial-pivoting tridiagonal LU), Ritz vectors
// y = V s are formed, orthonormalised (modified Gram-Schmidt, subspace preserving) and their
// residuals |C y - lambda y| are reported so the host can extend m if needed.
// Scores:  Y = (P - mean) V^T  (n x r), fp32.
#include "common.cuh"

namespace chr {

constexpr int kEigMaxM = 1024;   // Lanczos steps (and T dimension) upper bound
constexpr int kEigMaxR = 128;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += sh[w];   // fixed order
    return t;
}

// v0: deterministic pseudo-random start vector (the reference seeds ARPACK with uniform(-1,1), seed 42)
__global__ void lanczos_init_kernel(double* __restrict__ V, int64_t s) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) {
        uint64_t h = (uint64_t)(e + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        V[e] = u;
        acc += u * u;
    }
    const double nrm = sqrt(block_sum(acc, sh));
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) V[e] /= nrm;
}

// w = C v   (one warp per row, C symmetric row-major)
__global__ void symv_kernel(const double* __restrict__ C, const double* __restrict__ v, double* __restrict__ w, int64_t s) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= s) return;
    const double* __restrict__ c = C + row * s;
    double a = 0.0;
    for (int64_t e = lane; e < s; e += 32) a += c[e] * v[e];
    a = warp_sum(a);
    if (lane == 0) w[row] = a;
}

// h[i] = V[i] . w  for i < nv   (one block per vector)
__global__ void dots_kernel(const double* __restrict__ V, const double* __restrict__ w, int64_t s, double* __restrict__ h) {
    __shared__ double sh[32];
    const double* __restrict__ v = V + (int64_t)blockIdx.x * s;
    double a = 0.0;
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) a += v[e] * w[e];
    a = block_sum(a, sh);
    if (threadIdx.x == 0) h[blockIdx.x] = a;
}

// w -= sum_i h[i] V[i];  hacc[i] += h[i]
__global__ void project_out_kernel(const double* __restrict__ V, int nv, int64_t s, const double* __restrict__ h,
                                   double* __restrict__ w, double* __restrict__ hacc) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < s) {
        double a = w[e];
        for (int i = 0; i < nv; ++i) a -= h[i] * V[(int64_t)i * s + e];
        w[e] = a;
    }
    if (blockIdx.x == 0 && hacc)
        for (int i = threadIdx.x; i < nv; i += blockDim.x) hacc[i] += h[i];
}

// beta = |w|;  V[j+1] = w / beta;  alpha[j] = hacc[j]  (single block)
__global__ void lanczos_finish_step_kernel(double* __restrict__ V, int j, int64_t s, const double* __restrict__ w,
                                           double* __restrict__ hacc, double* __restrict__ alpha, double* __restrict__ beta,
                                           int write_next) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) a += w[e] * w[e];
    const double nrm = sqrt(block_sum(a, sh));
    const double inv = nrm > 1e-300 ? 1.0 / nrm : 0.0;
    const double aj = hacc[j];
    __syncthreads();
    if (write_next)
        for (int64_t e = threadIdx.x; e < s; e += blockDim.x) V[(int64_t)(j + 1) * s + e] = w[e] * inv;
    __syncthreads();
    if (threadIdx.x == 0) { alpha[j] = aj; beta[j] = nrm; }
    for (int i = threadIdx.x; i <= j; i += blockDim.x) hacc[i] = 0.0;
}

// ---- tridiagonal T (alpha[0..m), beta[0..m-1) off-diagonal): r largest eigenpairs -------------
__device__ int sturm_count_below(const double* a, const double* b, int m, double x, double pivmin) {
    int cnt = 0;
    double q = a[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0) ++cnt;
    for (int i = 1; i < m; ++i) {
        q = a[i] - x - b[i - 1] * b[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        if (q < 0) ++cnt;
    }
    return cnt;  // eigenvalues < x
}

// one thread per wanted eigenpair; S is [r][m] (row i = eigenvector of the i-th largest eigenvalue)
__global__ void tridiag_topr_kernel(const double* __restrict__ alpha, const double* __restrict__ beta, int m, int r,
                                    double* __restrict__ evals, double* __restrict__ S, double* __restrict__ scratch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    double lo = 1e300, hi = -1e300, tnorm = 0.0;
    for (int k = 0; k < m; ++k) {
        const double bl = k > 0 ? fabs(beta[k - 1]) : 0.0, br = k < m - 1 ? fabs(beta[k]) : 0.0;
        lo = fmin(lo, alpha[k] - bl - br);
        hi = fmax(hi, alpha[k] + bl + br);
        tnorm = fmax(tnorm, fabs(alpha[k]) + bl + br);
    }
    const double pivmin = 1e-290 * fmax(1.0, tnorm * tnorm);
    const int target = m - 1 - i;      // index (ascending) of the wanted eigenvalue: count(< lambda) == target
    double a = lo - 1e-12 * tnorm - 1e-300, b = hi + 1e-12 * tnorm + 1e-300;
    for (int it = 0; it < 120; ++it) {
        const double mid = 0.5 * (a + b);
        if (mid <= a || mid >= b) break;
        if (sturm_count_below(alpha, beta, m, mid, pivmin) <= target) a = mid; else b = mid;
    }
    const double lam = 0.5 * (a + b);
    evals[i] = lam;

    // inverse iteration on (T - lam I) with partial-pivoting LU (LAPACK dgttrf/dgtts2 structure)
    double* dl = scratch + (size_t)i * 5 * m;   // sub-diagonal multipliers
    double* d = dl + m;                          // diagonal of U
    double* du = d + m;                          // first super-diagonal of U
    double* du2 = du + m;                        // second super-diagonal of U
    double* y = du2 + m;                         // iterate
    const double eps = 2.2e-16 * fmax(tnorm, 1e-300);
    for (int k = 0; k < m; ++k) {
        d[k] = alpha[k] - lam;
        du[k] = k < m - 1 ? beta[k] : 0.0;
        dl[k] = k < m - 1 ? beta[k] : 0.0;
        du2[k] = 0.0;
    }
    // pivot flags are kept in the sign-free array `piv` packed into y temporarily
    int* piv = reinterpret_cast<int*>(S + (size_t)i * m);  // reuse the output row as int scratch (m ints <= m doubles)
    for (int k = 0; k < m - 1; ++k) {
        if (fabs(d[k]) >= fabs(dl[k])) {
            if (fabs(d[k]) < eps) d[k] = eps;
            const double f = dl[k] / d[k];
            dl[k] = f;
            d[k + 1] -= f * du[k];
            piv[k] = 0;
        } else {  // swap rows k and k+1
            const double f = d[k] / dl[k];
            d[k] = dl[k];
            dl[k] = f;
            const double t = du[k];
            du[k] = d[k + 1];
            d[k + 1] = t - f * du[k];
            if (k < m - 2) { du2[k] = du[k + 1]; du[k + 1] = -f * du[k + 1]; }
            piv[k] = 1;
        }
    }
    if (fabs(d[m - 1]) < eps) d[m - 1] = eps;
    for (int k = 0; k < m; ++k) {
        uint64_t h = (uint64_t)(k + 1 + 7919 * (i + 1)) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31;
        y[k] = 0.5 + (double)(h >> 40) * (1.0 / 16777216.0);
    }
    for (int iter = 0; iter < 4; ++iter) {
        // forward: L y = P b
        for (int k = 0; k < m - 1; ++k) {
            if (piv[k]) { const double t = y[k]; y[k] = y[k + 1]; y[k + 1] = t - dl[k] * y[k]; }
            else y[k + 1] -= dl[k] * y[k];
        }
        // backward: U x = y
        y[m - 1] /= d[m - 1];
        if (m > 1) y[m - 2] = (y[m - 2] - du[m - 2] * y[m - 1]) / d[m - 2];
        for (int k = m - 3; k >= 0; --k) y[k] = (y[k] - du[k] * y[k + 1] - du2[k] * y[k + 2]) / d[k];
        double nrm = 0.0, mx = 0.0;
        for (int k = 0; k < m; ++k) mx = fmax(mx, fabs(y[k]));
        const double sc = mx > 0 ? 1.0 / mx : 1.0;
        for (int k = 0; k < m; ++k) { y[k] *= sc; nrm += y[k] * y[k]; }
        nrm = 1.0 / sqrt(nrm);
        for (int k = 0; k < m; ++k) y[k] *= nrm;
    }
    double* out = S + (size_t)i * m;   // piv scratch is dead now
    for (int k = 0; k < m; ++k) out[k] = y[k];
}

// Y[i][e] = sum_j S[i][j] V[j][e]
__global__ void ritz_vectors_kernel(const double* __restrict__ S, const double* __restrict__ V, int m, int r, int64_t s,
                                    double* __restrict__ Y) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (e >= s) return;
    const double* __restrict__ si = S + (size_t)i * m;
    double a = 0.0;
    for (int j = 0; j < m; ++j) a += si[j] * V[(int64_t)j * s + e];
    Y[(int64_t)i * s + e] = a;
}

// modified Gram-Schmidt over the r Ritz vectors (single block; subspace preserving), twice
__global__ void mgs_kernel(double* __restrict__ Y, int r, int64_t s) {
    __shared__ double sh[32];
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < r; ++i) {
            double* yi = Y + (int64_t)i * s;
            for (int j = 0; j < i; ++j) {
                const double* yj = Y + (int64_t)j * s;
                double a = 0.0;
                for (int64_t e = threadIdx.x; e < s; e += blockDim.x) a += yi[e] * yj[e];
                a = block_sum(a, sh);
                for (int64_t e = threadIdx.x; e < s; e += blockDim.x) yi[e] -= a * yj[e];
                __syncthreads();
            }
            double a = 0.0;
            for (int64_t e = threadIdx.x; e < s; e += blockDim.x) a += yi[e] * yi[e];
            a = 1.0 / sqrt(block_sum(a, sh));
            for (int64_t e = threadIdx.x; e < s; e += blockDim.x) yi[e] *= a;
            __syncthreads();
        }
}

// lam[i] = y_i' C y_i (Rayleigh quotient) and res[i] = |C y_i - lam y_i|   (one block per vector)
__global__ void rayleigh_residual_kernel(const double* __restrict__ C, const double* __restrict__ Y, int64_t s,
                                         double* __restrict__ lam, double* __restrict__ res, double* __restrict__ cy_scratch) {
    __shared__ double sh[32];
    const int i = blockIdx.x;
    const double* __restrict__ y = Y + (int64_t)i * s;
    double* cy = cy_scratch + (int64_t)i * s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t row = warp; row < s; row += nw) {
        const double* __restrict__ c = C + row * s;
        double a = 0.0;
        for (int64_t e = lane; e < s; e += 32) a += c[e] * y[e];
        a = warp_sum(a);
        if (lane == 0) cy[row] = a;
    }
    __syncthreads();
    double a = 0.0;
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) a += y[e] * cy[e];
    const double l = block_sum(a, sh);
    a = 0.0;
    for (int64_t e = threadIdx.x; e < s; e += blockDim.x) { const double d = cy[e] - l * y[e]; a += d * d; }
    a = block_sum(a, sh);
    if (threadIdx.x == 0) { lam[i] = l; res[i] = sqrt(a); }
}

// ---- scores:  Y[i][c] = sum_g (P[i][g] - mean[g]) V[c][g] ------------------------------------
// skinny GEMM on CUDA cores (r <= 50 columns): a CTA owns 64 spots, walks the genes in chunks of 32
// (centred Z tile and V^T tile in shared memory), every thread accumulates 8 components of one spot.
// One pass over P per 32 components: HBM-bound (4 N S bytes).
constexpr int kScRows = 64, kScK = 32, kScComp = 32;
__global__ void __launch_bounds__(256)
pca_scores_kernel(const float* __restrict__ P, int64_t n, int64_t s, int64_t ld, const float* __restrict__ meanf,
                  const float* __restrict__ Vt /* [s][ncpad] */, int ncpad, int r, float* __restrict__ Y, int64_t ldy) {
    __shared__ float Zs[kScRows][kScK + 1];
    __shared__ __align__(16) float Vs[kScK][kScComp];
    const int row_l = threadIdx.x & 63, cgrp = threadIdx.x >> 6;
    const int64_t i0 = (int64_t)blockIdx.x * kScRows;
    for (int c0 = 0; c0 < r; c0 += kScComp) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int64_t g0 = 0; g0 < s; g0 += kScK) {
            for (int e = threadIdx.x; e < kScRows * kScK; e += 256) {
                const int rr = e >> 5, kk = e & 31;
                const int64_t i = i0 + rr, g = g0 + kk;
                Zs[rr][kk] = (i < n && g < s) ? __ldg(P + i * ld + g) - __ldg(meanf + g) : 0.f;
            }
            for (int e = threadIdx.x; e < kScK * kScComp; e += 256) {
                const int kk = e >> 5, c = e & 31;
                const int64_t g = g0 + kk;
                Vs[kk][c] = (g < s && c0 + c < ncpad) ? __ldg(Vt + g * ncpad + c0 + c) : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < kScK; ++k) {
                const float z = Zs[row_l][k];
                const float4 v0 = *reinterpret_cast<const float4*>(&Vs[k][cgrp * 8]);
                const float4 v1 = *reinterpret_cast<const float4*>(&Vs[k][cgrp * 8 + 4]);
                acc[0] = fmaf(z, v0.x, acc[0]); acc[1] = fmaf(z, v0.y, acc[1]);
                acc[2] = fmaf(z, v0.z, acc[2]); acc[3] = fmaf(z, v0.w, acc[3]);
                acc[4] = fmaf(z, v1.x, acc[4]); acc[5] = fmaf(z, v1.y, acc[5]);
                acc[6] = fmaf(z, v1.z, acc[6]); acc[7] = fmaf(z, v1.w, acc[7]);
            }
            __syncthreads();
        }
        const int64_t i = i0 + row_l;
        if (i < n)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + cgrp * 8 + j;
                if (c < r) Y[i * ldy + c] = acc[j];
            }
    }
}

// Vt[g][c] = (float) V[c][g], zero padded to ncpad components
__global__ void transpose_v_kernel(const double* __restrict__ V, int r, int64_t s, int ncpad, float* __restrict__ Vt) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= s * ncpad) return;
    const int64_t g = e / ncpad;
    const int c = (int)(e % ncpad);
    Vt[e] = c < r ? (float)V[(int64_t)c * s + g] : 0.f;
}

__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ b, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) b[e] = (float)a[e];
}

struct EigPlan {
    int m;
    size_t off_V, off_w, off_h, off_hacc, off_alpha, off_beta, off_S, off_scr, off_cy, total;
};

static EigPlan eig_plan(int64_t s, int r, int m) {
    EigPlan p;
    p.m = m;
    size_t o = 0;
    auto take = [&](size_t b) { size_t x = o; o += round_up<size_t>(b, 256); return x; };
    p.off_V = take((size_t)(m + 1) * s * 8);
    p.off_w = take((size_t)s * 8);
    p.off_h = take((size_t)(m + 1) * 8);
    p.off_hacc = take((size_t)(m + 1) * 8);
    p.off_alpha = take((size_t)(m + 1) * 8);
    p.off_beta = take((size_t)(m + 1) * 8);
    p.off_S = take((size_t)r * m * 8);
    p.off_scr = take((size_t)r * 5 * m * 8);
    p.off_cy = take((size_t)r * s * 8);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

int chr_eigh_default_steps(int64_t s, int r) {
    // first attempt of the host's doubling loop (device.eigh_topk): 3r + 32 steps already bring every Ritz residual of the
    // cfg4 covariance to 1e-15 (10 ms); 6r + 64 took 22 ms for the same answer (scripts/eigh_steps_check.py)
    int64_t m = 3 * (int64_t)r + 32;
    if (m > s) m = s;
    if (m > chr::kEigMaxM) m = chr::kEigMaxM;
    return (int)m;
}

size_t chr_eigh_workspace_bytes(int64_t s, int r, int m) {
    if (s <= 0 || r <= 0 || m <= 0) return 1024;
    return chr::eig_plan(s, r, m).total + 1024;
}

int chr_eigh_topk_f64(const double* C, int64_t s, int r, int m, double* evals, double* evecs, double* residuals,
                      void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(C && evals && evecs && residuals && workspace, "chr_eigh_topk_f64: null pointer argument");
    CHR_REQUIRE(s > 0 && r > 0 && r <= s && r <= kEigMaxR, "chr_eigh_topk_f64: need 0 < r <= min(s, %d) (s=%lld r=%d)", kEigMaxR,
                (long long)s, r);
    CHR_REQUIRE(m >= r && m <= s && m <= kEigMaxM, "chr_eigh_topk_f64: need r <= m <= min(s, %d) (m=%d)", kEigMaxM, m);
    EigPlan p = eig_plan(s, r, m);
    CHR_REQUIRE(workspace_bytes >= p.total + 256, "chr_eigh_topk_f64: workspace too small (%zu < %zu)", workspace_bytes, p.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* V = (double*)(ws + p.off_V);
    double* w = (double*)(ws + p.off_w);
    double* h = (double*)(ws + p.off_h);
    double* hacc = (double*)(ws + p.off_hacc);
    double* alpha = (double*)(ws + p.off_alpha);
    double* beta = (double*)(ws + p.off_beta);
    double* S = (double*)(ws + p.off_S);
    double* scr = (double*)(ws + p.off_scr);
    double* cy = (double*)(ws + p.off_cy);

    CHR_CUDA(cudaMemsetAsync(hacc, 0, (size_t)(m + 1) * 8, st));
    lanczos_init_kernel<<<1, 1024, 0, st>>>(V, s);
    const unsigned eb = (unsigned)ceil_div<int64_t>(s, 256);
    const unsigned rowb = (unsigned)ceil_div<int64_t>(s * 32, 256);
    for (int j = 0; j < m; ++j) {
        symv_kernel<<<rowb, 256, 0, st>>>(C, V + (int64_t)j * s, w, s);
        for (int pass = 0; pass < 2; ++pass) {
            dots_kernel<<<j + 1, 256, 0, st>>>(V, w, s, h);
            project_out_kernel<<<eb, 256, 0, st>>>(V, j + 1, s, h, w, hacc);
        }
        lanczos_finish_step_kernel<<<1, 1024, 0, st>>>(V, j, s, w, hacc, alpha, beta, j + 1 < m ? 1 : 0);
    }
    CHR_LAUNCH_CHECK();
    tridiag_topr_kernel<<<(unsigned)ceil_div(r, 32), 32, 0, st>>>(alpha, beta, m, r, evals, S, scr);
    CHR_LAUNCH_CHECK();
    dim3 rg(eb, (unsigned)r);
    ritz_vectors_kernel<<<rg, 256, 0, st>>>(S, V, m, r, s, evecs);
    mgs_kernel<<<1, 1024, 0, st>>>(evecs, r, s);
    rayleigh_residual_kernel<<<r, 512, 0, st>>>(C, evecs, s, evals, residuals, cy);
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_pca_scores_f32(const float* P, int64_t n, int64_t s, int64_t ld, const double* mean, const double* V, int r,
                       float* Y, int64_t ldy, void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(P && mean && V && Y && workspace, "chr_pca_scores_f32: null pointer argument");
    CHR_REQUIRE(n > 0 && s > 0 && r > 0 && ld >= s && ldy >= r, "chr_pca_scores_f32: bad shape");
    const size_t need = round_up<size_t>((size_t)s * 4, 256) + (size_t)round_up<int64_t>(r, kScComp) * s * 4 + 256;
    CHR_REQUIRE(workspace_bytes >= need, "chr_pca_scores_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* meanf = (float*)ws;
    float* Vt = (float*)(ws + round_up<size_t>((size_t)s * 4, 256));
    const int ncpad = (int)round_up<int64_t>(r, kScComp);
    f64_to_f32_kernel<<<(unsigned)ceil_div<int64_t>(s, 256), 256, 0, st>>>(mean, meanf, s);
    transpose_v_kernel<<<(unsigned)ceil_div<int64_t>(s * ncpad, 256), 256, 0, st>>>(V, r, s, ncpad, Vt);
    pca_scores_kernel<<<(unsigned)ceil_div<int64_t>(n, kScRows), 256, 0, st>>>(P, n, s, ld, meanf, Vt, ncpad, r, Y, ldy);
    CHR_LAUNCH_CHECK();
    return 0;
}

size_t chr_pca_scores_workspace_bytes(int64_t s, int r) {
    return chr::round_up<size_t>((size_t)(s > 0 ? s : 1) * 4, 256) + (size_t)chr::round_up<int64_t>(r > 0 ? r : 1, 32) * (size_t)(s > 0 ? s : 1) * 4 + 512;
}

}  // extern "C"
