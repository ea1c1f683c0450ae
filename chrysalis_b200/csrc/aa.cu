// K9-K11 - archetypal analysis: alternating penalised non-negative least squares.
//
// Replaces  archetypes.AA(n_archetypes, n_init=3, max_iter, tol=0.001, random_state=42).fit(X)
// (/root/reference/chrysalis/core.py:186-191; archetypes==0.4.2, pyproject.toml:24 - source not in
// the reference tree, algorithm restated in oracle/aa.py and SURVEY.md App. A.6).  One iteration:
//     alphas = optimize(X, Z)        N NNLS problems  min || [Z,M]^T a - [x_i,M] ||, a >= 0
//     Z      = pinv(alphas) X        normal equations A^T A (A x A), A^T X (A x d)
//     betas  = optimize(Z, X)        A NNLS problems with N unknowns each (<= d+1 non-zeros)
//     Z      = betas X ;  rss = || X - alphas Z ||_F^2
// with rows of alphas / betas renormalised to sum 1.  All arithmetic is fp64 (the reference
// upcasts to float64 inside numpy / scipy.optimize.nnls).
//
// NNLS is Lawson-Hanson's active-set method on the normal equations (scipy.optimize.nnls is the same method with QR
// updates; the minimisers are unique, so any correct active-set path gives the same coefficients).
//   alpha step  (N small problems sharing G = Z Z^T + M^2): one thread per spot.  Up to 16 archetypes a branch-free
//               kernel solves every passive set as a masked full-size Cholesky system so that the lanes of a warp never
//               diverge (aa_alpha_fixed_kernel); more archetypes take the index-list kernel (aa_alpha_kernel).  Both
//               warm-start from the previous coefficients and leave the CTA's partial A^T A | A^T X behind.
//   pinv step   one CTA: fixed-order reduction of the partials, Jacobi eigen-decomposition, numpy's cut-off.
//   beta step   (A problems with N unknowns each): every archetype advances one active-set step per sweep over X -
//               gradient = skinny product + arg-max (aa_beta_grad_*), then one warp per archetype refits its passive
//               set in shared memory (beta_refit_warp).  Passive sets are warm-started from the previous iteration.
#include "common.cuh"

namespace chr {

constexpr int kAaMaxA = 64;     // archetypes
constexpr int kAaMaxD = 64;     // embedding dimensions
constexpr int kAaMaxP = kAaMaxD + 2;   // passive-set bound for the beta problems (rank of [X, M] <= d + 1)

// measurement aid (scripts/aa_breakdown.py): [0] spots solved by warm alpha steps, [1] of those through the general
// active-set path, [2] warps with at least one such spot, [3] warps
#ifdef CHR_AA_STATS
__device__ unsigned long long g_aa_stats[4];   // measurement builds only (scripts/build_variant.sh aastats "-DCHR_AA_STATS" aa)
#endif

// ---- small dense helpers (per thread, column-major n x n in `a`, n <= 66) ----------------------
// Cholesky solve of the SPD system  H s = r  restricted to the index list `idx` (size np) of a
// full matrix `full` with leading dimension ld.  L is scratch (np*(np+1)/2).  Returns false if not SPD.
__device__ bool chol_solve_subset(const double* __restrict__ full, int ld, const int* idx, int np, const double* rhs_full,
                                  double* L, double* s) {
    for (int i = 0; i < np; ++i) {
        for (int j = 0; j <= i; ++j) {
            double v = full[idx[i] * ld + idx[j]];
            for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            if (i == j) {
                if (!(v > 0.0)) return false;
                L[i * (i + 1) / 2 + i] = sqrt(v);
            } else {
                L[i * (i + 1) / 2 + j] = v / L[j * (j + 1) / 2 + j];
            }
        }
    }
    for (int i = 0; i < np; ++i) {
        double v = rhs_full[idx[i]];
        for (int k = 0; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * s[k];
        s[i] = v / L[i * (i + 1) / 2 + i];
    }
    for (int i = np - 1; i >= 0; --i) {
        double v = s[i];
        for (int k = i + 1; k < np; ++k) v -= L[k * (k + 1) / 2 + i] * s[k];
        s[i] = v / L[i * (i + 1) / 2 + i];
    }
    return true;
}

// Lawson-Hanson NNLS on normal equations  min 1/2 a'Ga - c'a, a >= 0  (G: na x na, ld = na).
// x: in/out (na).  warm: x holds a feasible starting point (the previous iteration's coefficients); its
// support seeds the passive set, which usually is already the optimal one - the minimiser is unique (G is
// positive definite), so the result does not depend on the starting point.  Ties in the arg-max go to the
// lowest index.
template <int MAXA>
__device__ void nnls_normal(const double* __restrict__ G, const double* c, int na, double* x, bool warm) {
    int pidx[MAXA];
    double s[MAXA], w[MAXA], L[MAXA * (MAXA + 1) / 2];
    bool inP[MAXA];
    int np = 0;
    double cmax = 0.0;
    for (int j = 0; j < na; ++j) {
        inP[j] = warm && x[j] > 0.0;
        if (inP[j]) pidx[np++] = j; else x[j] = 0.0;
        cmax = fmax(cmax, fabs(c[j]));
    }
    const double tol = 1e-12 * fmax(cmax, 1e-300);
    // least squares on the passive set, stepping back to the feasible region while a coefficient turns non-positive
    auto settle = [&](bool added) {
        for (int inner = 0; inner < 3 * na && np > 0; ++inner) {
            if (!chol_solve_subset(G, na, pidx, np, c, L, s)) {   // numerically dependent column: drop the newest
                inP[pidx[np - 1]] = false;
                x[pidx[np - 1]] = 0.0;
                --np;
                if (added) return;
                continue;
            }
            double smin = 1e300;
            for (int i = 0; i < np; ++i) smin = fmin(smin, s[i]);
            if (smin > 0.0) {
                for (int i = 0; i < np; ++i) x[pidx[i]] = s[i];
                return;
            }
            double alpha = 1e300;
            for (int i = 0; i < np; ++i)
                if (s[i] <= 0.0) alpha = fmin(alpha, x[pidx[i]] / (x[pidx[i]] - s[i]));
            for (int i = 0; i < np; ++i) x[pidx[i]] += alpha * (s[i] - x[pidx[i]]);
            int q = 0;
            for (int i = 0; i < np; ++i) {
                if (x[pidx[i]] <= 1e-15 * fmax(1.0, cmax) && s[i] <= 0.0) { x[pidx[i]] = 0.0; inP[pidx[i]] = false; }
                else pidx[q++] = pidx[i];
            }
            np = q;
        }
    };
    if (np > 0) settle(false);
    for (int outer = 0; outer < 3 * na; ++outer) {
        for (int j = 0; j < na; ++j) {
            double v = c[j];
            for (int i = 0; i < np; ++i) v -= G[j * na + pidx[i]] * x[pidx[i]];
            w[j] = v;
        }
        int jbest = -1;
        double wbest = tol;
        for (int j = 0; j < na; ++j)
            if (!inP[j] && w[j] > wbest) { wbest = w[j]; jbest = j; }
        if (jbest < 0) break;
        inP[jbest] = true;
        pidx[np++] = jbest;
        settle(true);
    }
}

// Fast path of the warm-started alpha step: when the previous coefficients of a spot have at most kFastP
// non-zeros, the least-squares solve on that support runs entirely in registers (unrolled Cholesky, indices by
// __fns); if the solution is positive and every other archetype has a non-positive gradient it IS the NNLS
// minimiser (KKT; unique because G is positive definite) and the general active-set code is skipped.
// Returns false when the spot has to take the general path.
constexpr int kFastP = 6;
__device__ __forceinline__ bool nnls_warm_small(const double* __restrict__ G, const double* __restrict__ c, int na, unsigned mask,
                                                double tol, int* idx, double* sol) {
    const int np = __popc(mask);
    if (np == 0 || np > kFastP) return false;
    double L[kFastP][kFastP], rhs[kFastP];
#pragma unroll
    for (int p = 0; p < kFastP; ++p) {
        idx[p] = p < np ? (int)__fns(mask, 0, p + 1) : 0;
        rhs[p] = c[idx[p]];
    }
    // Cholesky of G[idx, idx] (rows/cols >= np are skipped by the guards)
#pragma unroll
    for (int i = 0; i < kFastP; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            if (i < np) {
                double v = G[idx[i] * na + idx[j]];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
                if (i == j) {
                    if (!(v > 0.0)) return false;
                    L[i][i] = sqrt(v);
                } else {
                    L[i][j] = v / L[j][j];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kFastP; ++i) {
        if (i < np) {
            double v = rhs[i];
#pragma unroll
            for (int k = 0; k < i; ++k) v -= L[i][k] * sol[k];
            sol[i] = v / L[i][i];
        }
    }
#pragma unroll
    for (int i = kFastP - 1; i >= 0; --i) {
        if (i < np) {
            double v = sol[i];
#pragma unroll
            for (int k = i + 1; k < kFastP; ++k)
                if (k < np) v -= L[k][i] * sol[k];
            sol[i] = v / L[i][i];
            if (!(sol[i] > 0.0)) return false;
        }
    }
    // KKT: no archetype outside the support may have a positive gradient
    for (int j = 0; j < na; ++j) {
        if ((mask >> j) & 1u) continue;
        double w = c[j];
#pragma unroll
        for (int p = 0; p < kFastP; ++p)
            if (p < np) w -= G[j * na + idx[p]] * sol[p];
        if (w > tol) return false;
    }
    return true;
}

// ---- alpha step: one thread per spot ---------------------------------------------------------------
// also accumulates the CTA's partial  A^T A (na x na)  and  A^T X (na x d)  for the pinv step
template <int MAXA>
__global__ void __launch_bounds__(128)
aa_alpha_kernel(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ Z, int na, double M, int warm,
                double* __restrict__ alphas, double* __restrict__ part /*[grid][na*na + na*d]*/) {
    extern __shared__ double sh[];
    double* G = sh;                       // na x na
    double* Zs = G + na * na;             // na x d
    double* acc = Zs + na * d;            // na*na + na*d  CTA accumulators
    const int nacc = na * na + na * d;
    for (int e = threadIdx.x; e < na * d; e += blockDim.x) Zs[e] = Z[e];
    for (int e = threadIdx.x; e < nacc; e += blockDim.x) acc[e] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) {
        const int a = e / na, b = e % na;
        double v = M * M;
        for (int t = 0; t < d; ++t) v += Zs[a * d + t] * Zs[b * d + t];
        G[e] = v;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double a_out[MAXA];
    bool valid = i < n;
    if (valid) {
        double c[MAXA];
        const double* __restrict__ x = X + i * d;
        for (int a = 0; a < na; ++a) {
            double v = M * M;
            for (int t = 0; t < d; ++t) v += Zs[a * d + t] * x[t];
            c[a] = v;
        }
        unsigned mask = 0;
        double cmax = 0.0;
        if (warm)
            for (int a = 0; a < na; ++a) {
                a_out[a] = alphas[i * na + a];
                if (a < 32 && a_out[a] > 0.0) mask |= 1u << a;
                cmax = fmax(cmax, fabs(c[a]));
            }
        int idx[kFastP];
        double sol[kFastP];
        const bool fast = warm && na <= 32 && nnls_warm_small(G, c, na, mask, 1e-12 * fmax(cmax, 1e-300), idx, sol);
#ifdef CHR_AA_STATS
        if (warm) {
            const unsigned act = __activemask(), slow = __ballot_sync(act, !fast);
            if ((threadIdx.x & 31) == __ffs(act) - 1) {
                atomicAdd(&g_aa_stats[0], (unsigned long long)__popc(act));
                atomicAdd(&g_aa_stats[1], (unsigned long long)__popc(slow));
                atomicAdd(&g_aa_stats[2], slow ? 1ull : 0ull);
                atomicAdd(&g_aa_stats[3], 1ull);
            }
        }
#endif
        if (fast) {
            for (int a = 0; a < na; ++a) a_out[a] = 0.0;
            const int np = __popc(mask);
#pragma unroll
            for (int p = 0; p < kFastP; ++p)
                if (p < np) a_out[idx[p]] = sol[p];
        } else {
            nnls_normal<MAXA>(G, c, na, a_out, warm != 0);
        }
        double sum = 0.0;
        for (int a = 0; a < na; ++a) { a_out[a] = fmax(a_out[a], 0.0); sum += a_out[a]; }
        const double inv = 1.0 / sum;
        for (int a = 0; a < na; ++a) { a_out[a] *= inv; alphas[i * na + a] = a_out[a]; }
    }
    // CTA-level accumulation in fixed order: the CTA's alpha rows go to shared memory, then every
    // thread owns accumulator entries and sums them over the spots 0..blockDim-1 in order
    double* As = acc + nacc;              // blockDim x na
    for (int a = 0; a < na; ++a) As[threadIdx.x * na + a] = valid ? a_out[a] : 0.0;
    __syncthreads();
    const int64_t i_base = (int64_t)blockIdx.x * blockDim.x;
    const int nvalid = (int)((n - i_base) < (int64_t)blockDim.x ? (n - i_base) : (int64_t)blockDim.x);
    for (int e = threadIdx.x; e < nacc; e += blockDim.x) {
        double v = 0.0;
        if (e < na * na) {
            const int a = e / na, b = e % na;
            for (int t = 0; t < nvalid; ++t) v += As[t * na + a] * As[t * na + b];
        } else {
            const int a = (e - na * na) / d, q = (e - na * na) % d;
            const double* __restrict__ xq = X + i_base * d + q;
            for (int t = 0; t < nvalid; ++t) v += As[t * na + a] * xq[(int64_t)t * d];
        }
        acc[e] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nacc; e += blockDim.x) part[(size_t)blockIdx.x * nacc + e] = acc[e];
}

// stage `nrows` contiguous rows of `d` doubles into shared-memory rows of stride `ld`: flat coalesced sweep with 8-byte
// cp.async (everything in flight at once, no registers; (row, column) advance incrementally - no division in the loop).
// The caller waits with stage_wait() and a CTA barrier.
__device__ __forceinline__ void stage_rows(double* __restrict__ dst, int ld, const double* __restrict__ src, int d, int nrows) {
    const int T = blockDim.x, total = nrows * d;
    const int q = T / d, rm = T - q * d;
    int row = threadIdx.x / d, col = threadIdx.x - row * d;
    for (int e = threadIdx.x; e < total; e += T) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + row * ld + col)), "l"(src + e)
                     : "memory");
        col += rm;
        row += q;
        if (col >= d) { col -= d; ++row; }
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- alpha step for n_archetypes <= 16: branch-free masked Cholesky, one thread per spot --------------------
// The active-set iteration of a spot is written so that every lane of a warp executes the same instruction
// stream whatever its passive set P is: the least-squares problem on P is solved as the FULL NA x NA system
//     G'_ij = G_ij  (i, j in P),   delta_ij otherwise;      rhs'_i = c_i (i in P), 0 otherwise
// (NA = n_archetypes padded to a multiple of 4, padding rows are never passive), with a compile-time unrolled
// Cholesky: up to 12 archetypes the matrix and its factor stay in registers, 13 - 16 keep the factor in shared memory as
// L[entry][thread] (conflict-free) with the row being eliminated in registers.  One loop iteration = one solve, then either "accept + KKT test + add the arg-max archetype" or
// "step back to the feasible region + drop" (Lawson-Hanson, same tolerances as nnls_normal); lanes that have
// converged idle.  A warm start (previous coefficients) usually needs 1-3 solves.  The minimiser is unique, so
// this and nnls_normal agree to rounding.
template <int NA>
struct AlphaFixed {
    static constexpr int kTri = NA * (NA + 1) / 2;
    static constexpr int kThreads = 128;

    // row-oriented variant (factor rows re-read from shared memory): used when the matrix does not fit the register file (NA = 16)
        static __device__ __forceinline__ void solve_rows(const double* __restrict__ G, unsigned& P, const double (&c)[NA], double* __restrict__ Lt,
                                                 double (&s)[NA]) {
        double dinv[NA], y[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const bool mi = (P >> i) & 1u;
            double row[NA];
#pragma unroll
            for (int j = 0; j < i; ++j) {
                double v = (mi && ((P >> j) & 1u)) ? G[i * NA + j] : 0.0;
#pragma unroll
                for (int k = 0; k < j; ++k) v -= row[k] * Lt[(j * (j + 1) / 2 + k) * kThreads];
                row[j] = v * dinv[j];
            }
            const double gii = G[i * NA + i];
            double v = mi ? gii : 1.0;
#pragma unroll
            for (int k = 0; k < i; ++k) v -= row[k] * row[k];
            const bool ok = v > 1e-13 * gii || !mi;           // numerically dependent archetype: leave it out
            if (!ok) P &= ~(1u << i);
            v = ok ? v : 1.0;
            dinv[i] = rsqrt(v);
            double r = (mi && ok) ? c[i] : 0.0;
#pragma unroll
            for (int j = 0; j < i; ++j) {
                row[j] = ok ? row[j] : 0.0;
                Lt[(i * (i + 1) / 2 + j) * kThreads] = row[j];
                r -= row[j] * y[j];
            }
            y[i] = r * dinv[i];
        }
#pragma unroll
        for (int i = NA - 1; i >= 0; --i) {
            double v = y[i];
#pragma unroll
            for (int k = i + 1; k < NA; ++k) v -= Lt[(k * (k + 1) / 2 + i) * kThreads] * s[k];
            s[i] = v * dinv[i];
        }
    }

    // solve on the passive set `P` (bit mask); returns s (0 outside P); indices whose pivot collapses are dropped from P.
    // Right-looking Cholesky with the (masked) matrix in registers: column j is scaled IN PLACE (the factor replaces the matrix
    // entry by entry; the back substitution reads it from the same registers - no shared-memory round trip) and subtracted from
    // the trailing entries; the forward substitution is fused.
    // Every entry sees the subtractions in the order k = 0, 1, ... of the inner-product form.
    static __device__ __forceinline__ void solve(const double* __restrict__ G, unsigned& P, const double (&c)[NA], double* __restrict__ Lt,
                                                 double (&s)[NA]) {
        if constexpr (NA > 12) {
            solve_rows(G, P, c, Lt, s);
            return;
        }
        double A[kTri], y[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const bool mi = (P >> i) & 1u;
#pragma unroll
            for (int j = 0; j < i; ++j) A[i * (i + 1) / 2 + j] = (mi && ((P >> j) & 1u)) ? G[i * NA + j] : 0.0;
            A[i * (i + 1) / 2 + i] = mi ? G[i * NA + i] : 1.0;
            y[i] = mi ? c[i] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const bool mj = (P >> j) & 1u;
            double v = A[j * (j + 1) / 2 + j];
            const bool ok = v > 1e-13 * G[j * NA + j] || !mj;         // numerically dependent archetype: leave it out
            if (!ok) P &= ~(1u << j);
            v = ok ? v : 1.0;
            const double dj = rsqrt(v);
            A[j * (j + 1) / 2 + j] = dj;                                // the diagonal slot keeps 1 / L_jj
            y[j] = ok ? y[j] * dj : 0.0;
#pragma unroll
            for (int i = j + 1; i < NA; ++i) {
                const double c = ok ? A[i * (i + 1) / 2 + j] * dj : 0.0;
                A[i * (i + 1) / 2 + j] = c;                             // the factor replaces the matrix entry by entry (registers)
                y[i] -= c * y[j];
            }
#pragma unroll
            for (int i = j + 1; i < NA; ++i)
#pragma unroll
                for (int k = j + 1; k <= i; ++k) A[i * (i + 1) / 2 + k] -= A[i * (i + 1) / 2 + j] * A[k * (k + 1) / 2 + j];
        }
#pragma unroll
        for (int i = NA - 1; i >= 0; --i) {
            double v = y[i];
#pragma unroll
            for (int k = i + 1; k < NA; ++k) v -= A[k * (k + 1) / 2 + i] * s[k];
            s[i] = v * A[i * (i + 1) / 2 + i];
        }
    }
};

template <int NA>
__global__ void __launch_bounds__(128)
aa_alpha_fixed_kernel(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ Z, int na, double M, int warm,
                      double* __restrict__ alphas, double* __restrict__ part /*[grid][na*na + na*d]*/, int region_doubles) {
    using AF = AlphaFixed<NA>;
    constexpr int T = AF::kThreads;
    extern __shared__ __align__(16) double sh[];
    double* G = sh;                        // NA x NA, identity on the padding
    double* Zs = G + NA * NA;              // d x NA, transposed and zero-padded: the archetypes of one coordinate are contiguous
    double* region = Zs + NA * d;          // staged X rows -> Cholesky factors -> [alphas | X] rows of the CTA
    const int tid = threadIdx.x;
    const int nacc = na * na + na * d;
    const int dpad = d | 1;                // odd row stride: conflict-free per-thread rows
    double* A0 = region + region_doubles;  // T x (na | 1): the previous coefficients of the CTA's spots (warm start)
    const int64_t i_base = (int64_t)blockIdx.x * T;
    const int nvalid = (int)((n - i_base) < (int64_t)T ? (n - i_base) : (int64_t)T);
    for (int e = tid; e < NA * d; e += T) {
        const int t = e / NA, a = e - t * NA;
        Zs[e] = a < na ? Z[a * d + t] : 0.0;
    }
    stage_rows(region, dpad, X + i_base * d, d, nvalid);   // the CTA's rows of X: one contiguous block
    if (warm) stage_rows(A0, na | 1, alphas + i_base * na, na, nvalid);
    stage_wait();
    __syncthreads();
    for (int e = tid; e < NA * NA; e += T) {
        const int a = e / NA, b = e % NA;
        double v = a == b ? 1.0 : 0.0;
        if (a < na && b < na) {
            v = M * M;
            for (int t = 0; t < d; ++t) v += Zs[t * NA + a] * Zs[t * NA + b];
        }
        G[e] = v;
    }
    const bool valid = tid < nvalid;
    double c[NA], x[NA], s[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) { c[a] = a < na ? M * M : 0.0; x[a] = 0.0; s[a] = 0.0; }
    if (valid) {
        const double* __restrict__ xr = region + tid * dpad;
        for (int t = 0; t < d; ++t) {
            const double xv = xr[t];
            const double2* __restrict__ z2 = reinterpret_cast<const double2*>(Zs + t * NA);
#pragma unroll
            for (int h = 0; h < NA / 2; ++h) {
                const double2 z = z2[h];                               // padding archetypes: z = 0
                c[2 * h] += z.x * xv;
                c[2 * h + 1] += z.y * xv;
            }
        }
    }
    unsigned P = 0;
    double cmax = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a) cmax = fmax(cmax, fabs(c[a]));
    if (valid && warm) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
            if (a < na) {
                x[a] = A0[tid * (na | 1) + a];
                if (x[a] > 0.0) P |= 1u << a; else x[a] = 0.0;
            }
    }
    __syncthreads();                       // G complete; every thread is done with its staged row (the factors go there)
    const double tol = 1e-12 * fmax(cmax, 1e-300), tiny = 1e-15 * fmax(1.0, cmax);
    double* Lt = region + tid;
    bool active = valid;
#ifdef CHR_AA_STATS
    int my_solves = 0, warp_solves = 0;
#endif
    for (int it = 0; it < 8 * NA && __any_sync(0xffffffffu, active); ++it) {
#ifdef CHR_AA_STATS
        my_solves += active ? 1 : 0;
        ++warp_solves;
#endif
        unsigned Pn = P;
        AF::solve(G, Pn, c, Lt, s);
        if (active) {
            if (Pn != P) {                 // collapsed pivots: those archetypes leave the passive set
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if (((P ^ Pn) >> a) & 1u) x[a] = 0.0;
                P = Pn;
            }
            bool feas = true;
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if ((P >> a) & 1u) feas = feas && s[a] > 0.0;
            if (feas) {
#pragma unroll
                for (int a = 0; a < NA; ++a) x[a] = ((P >> a) & 1u) ? s[a] : 0.0;
            } else {
                double alpha = 1e300;
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if (((P >> a) & 1u) && s[a] <= 0.0) alpha = fmin(alpha, x[a] / (x[a] - s[a]));
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if ((P >> a) & 1u) {
                        x[a] += alpha * (s[a] - x[a]);
                        if (x[a] <= tiny && s[a] <= 0.0) { x[a] = 0.0; P &= ~(1u << a); }
                    }
            }
            if (feas) {                    // KKT: the largest gradient outside the passive set joins it, lowest index on ties
                int jbest = -1;
                double wbest = tol;
#pragma unroll
                for (int j = 0; j < NA; ++j) {
                    double w = c[j];
#pragma unroll
                    for (int a = 0; a < NA; ++a) w -= G[j * NA + a] * x[a];
                    if (j < na && !((P >> j) & 1u) && w > wbest) { wbest = w; jbest = j; }
                }
                if (jbest < 0) active = false; else P |= 1u << jbest;
            }
        }
    }
#ifdef CHR_AA_STATS
    if (warm) {   // [0] spots, [1] solves the spots needed, [2] solves the warps executed (x 32 lanes), [3] warps
        atomicAdd(&g_aa_stats[0], valid ? 1ull : 0ull);
        atomicAdd(&g_aa_stats[1], (unsigned long long)my_solves);
        if ((tid & 31) == 0) { atomicAdd(&g_aa_stats[2], (unsigned long long)warp_solves); atomicAdd(&g_aa_stats[3], 1ull); }
    }
#endif
    double sum = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a) { x[a] = fmax(x[a], 0.0); sum += x[a]; }
    const double inv = 1.0 / sum;
    __syncthreads();                       // the factors are dead: the region becomes B = [alphas | X | 0] (row = spot)
    // rows padded to whole strips of 4 columns; the row stride is = 2 (mod 4) doubles: 16-byte aligned rows, and the
    // per-thread row writes below see 2-way instead of 4-way bank conflicts
    const int W = na + d, Wp4 = (W + 3) & ~3, S = Wp4 + 2;
#pragma unroll
    for (int a = 0; a < NA; ++a)
        if (a < na) region[tid * S + a] = valid ? x[a] * inv : 0.0;
    for (int c = W; c < Wp4; ++c) region[tid * S + c] = 0.0;
    stage_rows(region + na, S, X + i_base * d, d, nvalid);
    for (int r = nvalid + (tid >> 5); r < T; r += T >> 5)          // rows past the end of the matrix contribute zeros
        for (int t = tid & 31; t < d; t += 32) region[r * S + na + t] = 0.0;
    stage_wait();
    __syncthreads();
    {   // the CTA's alphas: one contiguous block of the output, written coalesced
        const int q = T / na, rm = T - q * na;
        int row = tid / na, col = tid - row * na;
        for (int e = tid; e < nvalid * na; e += T) {
            alphas[i_base * na + e] = region[row * S + col];
            col += rm;
            row += q;
            if (col >= na) { col -= na; ++row; }
        }
    }
    // partial  A^T [A | X]  of the CTA (na x W) as 4 x 4 register blocks: block row ab (archetypes 4ab .. 4ab+3) times
    // strip st (columns 4st .. 4st+3), 16 products for four 16-byte loads.  Strips entirely below the diagonal of A^T A
    // are skipped (their mirror images are written instead).  Warp w sums the spots 32w .. 32w+31, its lanes own the
    // blocks; the four partial sums of a block are added in warp order (fixed summation tree).
    const int nstrip = Wp4 >> 2;
    int nE = 0;
    for (int ab = 0; ab < NA / 4; ++ab) nE += nstrip - ab;
    double* red = region + T * S;          // [4 warps][nE][16]
    const int lane = tid & 31, warp = tid >> 5;
    for (int e = lane; e < nE; e += 32) {
        int ab = 0, rem = e;
        while (rem >= nstrip - ab) { rem -= nstrip - ab; ++ab; }
        const int st = ab + rem;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;
        const double* __restrict__ b = region + (warp * 32) * S;
        for (int t = 0; t < 32; ++t, b += S) {
            const double2 a01 = *reinterpret_cast<const double2*>(b + 4 * ab), a23 = *reinterpret_cast<const double2*>(b + 4 * ab + 2);
            const double2 c01 = *reinterpret_cast<const double2*>(b + 4 * st), c23 = *reinterpret_cast<const double2*>(b + 4 * st + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] += av[r] * cv[q];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) red[((size_t)warp * nE + e) * 16 + r * 4 + q] = acc[r][q];
    }
    __syncthreads();
    for (int o = tid; o < nE * 16; o += T) {
        const int e = o >> 4, r = (o >> 2) & 3, q = o & 3;
        int ab = 0, rem = e;
        while (rem >= nstrip - ab) { rem -= nstrip - ab; ++ab; }
        const int a = 4 * ab + r, col = 4 * (ab + rem) + q;
        if (a >= na || col >= W) continue;
        double v = red[o];
#pragma unroll
        for (int w = 1; w < T / 32; ++w) v += red[(size_t)w * nE * 16 + o];
        // same layout as aa_alpha_kernel: [na x na | na x d]
        const int dst = col < na ? a * na + col : na * na + a * d + (col - na);
        part[(size_t)blockIdx.x * nacc + dst] = v;
        if (col < na && (col >> 2) > (a >> 2)) part[(size_t)blockIdx.x * nacc + col * na + a] = v;   // skipped mirror strip
    }
}

static size_t aa_alpha_fixed_region_doubles(int NA, int na, int d) {
    const int W = na + d, Wp4 = (W + 3) & ~3, S = Wp4 + 2, nstrip = Wp4 >> 2;
    int nE = 0;
    for (int ab = 0; ab < NA / 4; ++ab) nE += nstrip - ab;
    const size_t tri = NA > 12 ? (size_t)NA * (NA + 1) / 2 * 128 : 0;              // Cholesky factors (NA <= 12 keeps them in registers)
    const size_t xs = (size_t)128 * (d | 1);                                       // staged rows
    const size_t b = (size_t)128 * S + (size_t)4 * nE * 16;                        // [alphas | X] rows + the warps' partial blocks
    size_t r = tri > xs ? tri : xs;
    return r > b ? r : b;
}

// fixed-order reduction of the per-CTA partials  A^T A | A^T X  in chunks of 64 CTAs (second level in aa_pinv_kernel)
constexpr int kAaRedChunk = 64;
__global__ void aa_partial_reduce_kernel(const double* __restrict__ part, int nparts, int nacc, double* __restrict__ part2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nacc) return;
    const int p0 = blockIdx.y * kAaRedChunk, p1 = p0 + kAaRedChunk < nparts ? p0 + kAaRedChunk : nparts;
    double a = 0.0;
    for (int p = p0; p < p1; ++p) a += part[(size_t)p * nacc + e];
    part2[(size_t)blockIdx.y * nacc + e] = a;
}

// ---- pinv step (single CTA): Zhat = pinv(alphas) X via eigen-decomposition of A^T A ---------------------
__global__ void aa_pinv_kernel(const double* __restrict__ part, int nparts, int na, int d, int64_t n, double* __restrict__ Zhat,
                               double* __restrict__ gram_out /* [na*na + na*d]: A^T A | A^T X, for the RSS identity */) {
    extern __shared__ double sh[];
    const int nacc = na * na + na * d;
    double* S = sh;                 // na x na  (A^T A, becomes diagonal)
    double* B = S + na * na;        // na x d   (A^T X)
    double* Vv = B + na * d;        // na x na  eigenvectors (columns)
    for (int e = threadIdx.x; e < nacc; e += blockDim.x) {
        double a = 0.0;
#pragma unroll 16
        for (int p = 0; p < nparts; ++p) a += part[(size_t)p * nacc + e];          // loads issue ahead, the sum stays in order
        if (e < na * na) S[e] = a; else B[e - na * na] = a;
        gram_out[e] = a;
    }
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) Vv[e] = (e / na == e % na) ? 1.0 : 0.0;
    __syncthreads();
    if (threadIdx.x < 32) {
        // Jacobi eigen-decomposition by one warp, round-robin ordering: a round holds na/2 rotations on disjoint index
        // pairs - they commute, so their angles are taken from the same S and their column and row updates run in parallel
        const int lane = threadIdx.x;
        const int m = (na + 1) & ~1;                       // players of the round-robin schedule (an odd na gets a bye)
        double* cs = Vv + na * na;                         // [m/2][2] cosine, sine
        int* pq = reinterpret_cast<int*>(cs + m);          // [m/2] p | q << 8, q >= na: bye
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0.0, diag = 0.0;
            for (int e = lane; e < na * na; e += 32) {
                const int p = e / na, q = e - p * na;
                const double v = S[e];
                if (p == q) diag += v * v; else if (q > p) off += v * v;
            }
            off = warp_sum(off);
            diag = warp_sum(diag);
            if (off <= 1e-30 * diag) break;                // butterfly sums: the same value in every lane
            for (int r = 0; r < m - 1; ++r) {
                if (lane < m / 2) {
                    int p = lane == 0 ? m - 1 : (r + lane) % (m - 1);
                    int q = lane == 0 ? r : (r - lane + (m - 1)) % (m - 1);
                    if (p > q) { const int t = p; p = q; q = t; }
                    double c = 1.0, sn = 0.0;
                    if (q < na) {
                        const double apq = S[p * na + q];
                        if (fabs(apq) >= 1e-300) {
                            const double theta = (S[q * na + q] - S[p * na + p]) / (2.0 * apq);
                            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                            c = 1.0 / sqrt(t * t + 1.0);
                            sn = t * c;
                        }
                    }
                    cs[2 * lane] = c;
                    cs[2 * lane + 1] = sn;
                    pq[lane] = p | (q << 8);
                }
                __syncwarp();
                for (int it = lane; it < na * (m / 2); it += 32) {          // columns p, q of S and V
                    const int pr = it / na, k = it - pr * na;
                    const int p = pq[pr] & 0xff, q = pq[pr] >> 8;
                    if (q >= na) continue;
                    const double c = cs[2 * pr], sn = cs[2 * pr + 1];
                    const double skp = S[k * na + p], skq = S[k * na + q];
                    S[k * na + p] = c * skp - sn * skq;
                    S[k * na + q] = sn * skp + c * skq;
                    const double vkp = Vv[k * na + p], vkq = Vv[k * na + q];
                    Vv[k * na + p] = c * vkp - sn * vkq;
                    Vv[k * na + q] = sn * vkp + c * vkq;
                }
                __syncwarp();
                for (int it = lane; it < na * (m / 2); it += 32) {          // rows p, q of S
                    const int pr = it / na, k = it - pr * na;
                    const int p = pq[pr] & 0xff, q = pq[pr] >> 8;
                    if (q >= na) continue;
                    const double c = cs[2 * pr], sn = cs[2 * pr + 1];
                    const double spk = S[p * na + k], sqk = S[q * na + k];
                    S[p * na + k] = c * spk - sn * sqk;
                    S[q * na + k] = sn * spk + c * sqk;
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // pinv(A) X = V diag(1/lambda) V^T (A^T X);  numpy.linalg.pinv drops singular values below
    // rtol * sigma_max with rtol = max(n, na) * eps, i.e. eigenvalues below (rtol^2) * lambda_max
    double lmax = 0.0;
    for (int p = 0; p < na; ++p) lmax = fmax(lmax, S[p * na + p]);
    const double rt = (double)(n > na ? n : na) * 2.220446049250313e-16;
    const double cut = fmax(rt * rt, 1e-14) * lmax;   // eigenvalues of A^T A are only resolved to ~1e-16 * lmax
    for (int e = threadIdx.x; e < na * d; e += blockDim.x) {
        const int a = e / d, q = e % d;
        double v = 0.0;
        for (int p = 0; p < na; ++p) {
            const double lam = S[p * na + p];
            if (lam <= cut) continue;
            double t = 0.0;
            for (int b = 0; b < na; ++b) t += Vv[b * na + p] * B[b * d + q];
            v += Vv[a * na + p] * t / lam;
        }
        Zhat[e] = v;
    }
}

// ---- beta step -------------------------------------------------------------------------------------
// state per archetype a: passive list pidx[a][<=kAaMaxP], coefficient b[a][.], count np[a], done[a];
// residual R[a][0..d] = [zhat_a, M] - sum_p b_p [x_p, M]
struct BetaState {
    int np[kAaMaxA];
    int done[kAaMaxA];
    int all_done;
    int iters;
};

// gradient pass:  w[i][a] = [x_i, M] . R[a];  per CTA arg-max over its spots for every archetype
__global__ void __launch_bounds__(256)
aa_beta_grad_kernel(const double* __restrict__ X, int64_t n, int d, double M, const double* __restrict__ R, int na,
                    const BetaState* __restrict__ st, const int* __restrict__ pidx, double* __restrict__ cand_w,
                    int64_t* __restrict__ cand_i) {
    if (st->all_done) return;
    extern __shared__ double sh[];
    double* Rs = sh;                                   // na x (d+1)
    double* red_w = Rs + na * (d + 1);                 // [warps]
    int64_t* red_i = reinterpret_cast<int64_t*>(red_w + 32);
    for (int e = threadIdx.x; e < na * (d + 1); e += blockDim.x) Rs[e] = R[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * 1024;
    const int64_t i1 = i0 + 1024 < n ? i0 + 1024 : n;
    for (int a = 0; a < na; ++a) {
        if (st->done[a]) {
            if (threadIdx.x == 0) { cand_w[(size_t)blockIdx.x * na + a] = -1e300; cand_i[(size_t)blockIdx.x * na + a] = -1; }
            continue;
        }
        const double* __restrict__ r = Rs + a * (d + 1);
        double bw = -1e300;
        int64_t bi = -1;
        for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            const double* __restrict__ x = X + i * d;
            double v = M * r[d];
            for (int t = 0; t < d; ++t) v += x[t] * r[t];
            if (v > bw) { bw = v; bi = i; }
        }
        // passive entries are excluded by the update kernel (their gradient is ~0); arg-max, lowest index wins ties
        for (int o = 16; o > 0; o >>= 1) {
            const double ow = __shfl_xor_sync(0xffffffffu, bw, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ow > bw || (ow == bw && oi >= 0 && (bi < 0 || oi < bi))) { bw = ow; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { red_w[warp] = bw; red_i[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < nw; ++w)
                if (red_w[w] > bw || (red_w[w] == bw && red_i[w] >= 0 && (bi < 0 || red_i[w] < bi))) { bw = red_w[w]; bi = red_i[w]; }
            cand_w[(size_t)blockIdx.x * na + a] = bw;
            cand_i[(size_t)blockIdx.x * na + a] = bi;
        }
    }
    (void)pidx;
}

// gradient pass for n_archetypes <= 16: persistent CTAs walk the spots in chunks of 256 rows, staged in shared memory
// (coalesced); every thread evaluates all archetypes for its row in one sweep, so X is read once per pass instead of once
// per archetype.  Same arithmetic per (row, archetype) as aa_beta_grad_kernel and ties go to the lowest index, so the
// winners are the same whatever the grid is.
constexpr int kGradChunk = 256, kGradThreads = 128;      // every thread owns 2 rows of a chunk: the residual rows are loaded once for both
template <int MAXA>
__global__ void __launch_bounds__(kGradThreads)
aa_beta_grad_fixed_kernel(const double* __restrict__ X, int64_t n, int d, double M, const double* __restrict__ R, int na,
                          const BetaState* __restrict__ st, double* __restrict__ cand_w, int64_t* __restrict__ cand_i) {
    if (st->all_done) return;
    extern __shared__ __align__(16) double sh[];
    const int dpad = d | 1;
    double* Rt = sh;                                   // (d+1) x MAXA, transposed: the archetypes of one coordinate are contiguous
    double* Xs = Rt + (d + 1) * MAXA;                  // kGradChunk x dpad
    double* red_w = Xs + kGradChunk * dpad;            // [4][MAXA]
    int64_t* red_i = reinterpret_cast<int64_t*>(red_w + 4 * MAXA);
    for (int e = threadIdx.x; e < (d + 1) * MAXA; e += kGradThreads) {
        const int t = e / MAXA, a = e % MAXA;
        Rt[e] = a < na ? R[a * (d + 1) + t] : 0.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nchunks = (n + kGradChunk - 1) / kGradChunk;
    double bw[MAXA];
    int64_t bi[MAXA];
#pragma unroll
    for (int a = 0; a < MAXA; ++a) { bw[a] = -1e300; bi[a] = -1; }
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t r0 = c * kGradChunk;
        const int nrows = (int)((n - r0) < kGradChunk ? (n - r0) : kGradChunk);
        __syncthreads();                               // previous chunk consumed (and Rt written, first time)
        stage_rows(Xs, dpad, X + r0 * d, d, nrows);
        stage_wait();
        __syncthreads();
        // rows tid and tid + 128 (a row past the end shadows row 0 and is not offered as a candidate)
        const int ra = threadIdx.x, rb = threadIdx.x + kGradThreads;
        const double* __restrict__ xa = Xs + (ra < nrows ? ra : 0) * dpad;
        const double* __restrict__ xb = Xs + (rb < nrows ? rb : 0) * dpad;
        double va[MAXA], vb[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) va[a] = vb[a] = M * Rt[d * MAXA + a];
        for (int t = 0; t < d; ++t) {
            const double xva = xa[t], xvb = xb[t];
            const double2* __restrict__ r2 = reinterpret_cast<const double2*>(Rt + t * MAXA);
#pragma unroll
            for (int h = 0; h < MAXA / 2; ++h) {
                const double2 r = r2[h];
                va[2 * h] += xva * r.x; va[2 * h + 1] += xva * r.y;
                vb[2 * h] += xvb * r.x; vb[2 * h + 1] += xvb * r.y;
            }
        }
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {               // rows ascend (ra before rb, chunks in order): ties keep the lowest index
            if (a < na && ra < nrows && va[a] > bw[a]) { bw[a] = va[a]; bi[a] = r0 + ra; }
            if (a < na && rb < nrows && vb[a] > bw[a]) { bw[a] = vb[a]; bi[a] = r0 + rb; }
        }
    }
    // arg-max per archetype over the CTA, lowest index wins ties
#pragma unroll
    for (int a = 0; a < MAXA; ++a) {
        if (a < na) {
            double w = bw[a];
            int64_t i = bi[a];
            for (int o = 16; o > 0; o >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, w, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, i, o);
                if (ow > w || (ow == w && oi >= 0 && (i < 0 || oi < i))) { w = ow; i = oi; }
            }
            if (lane == 0) { red_w[warp * MAXA + a] = w; red_i[warp * MAXA + a] = i; }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < na) {
        const int a = threadIdx.x;
        double w = red_w[a];
        int64_t i = red_i[a];
        for (int q = 1; q < kGradThreads / 32; ++q) {
            const double ow = red_w[q * MAXA + a];
            const int64_t oi = red_i[q * MAXA + a];
            if (ow > w || (ow == w && oi >= 0 && (i < 0 || oi < i))) { w = ow; i = oi; }
        }
        const bool done = st->done[a] != 0;
        cand_w[(size_t)blockIdx.x * na + a] = done ? -1e300 : w;
        cand_i[(size_t)blockIdx.x * na + a] = done ? -1 : i;
    }
}

// least squares of [zhat_a, M] on the passive points P of archetype a, by one warp: steps back to the feasible region
// while a coefficient turns non-positive, then writes the residual
//   R[a] = [zhat_a, M] - sum_p b_p [x_p, M]
// Shared memory of the warp (doubles): passive rows Xp [npmax][d|1], Gram / Cholesky factor [npmax][npmax|1], rhs, s.
// The Gram entries and the factor are computed entry by entry with the accumulation order of a serial solve (lanes own
// rows), the triangular solves sweep column by column.
__device__ inline int beta_npmax(int d) { return d + 3 < kAaMaxP ? d + 3 : kAaMaxP; }
static size_t beta_refit_smem_bytes(int d) {
    const int npmax = d + 3 < kAaMaxP ? d + 3 : kAaMaxP;
    return ((size_t)npmax * (d | 1) + (size_t)npmax * (npmax | 1) + 2 * (size_t)npmax) * 8;
}

__device__ void beta_refit_warp(const double* __restrict__ X, int d, double M, const double* __restrict__ Zhat, int a, BetaState* st,
                                int* P, double* b, double* __restrict__ R, double* sm) {
    const int lane = threadIdx.x & 31;
    const int npmax = beta_npmax(d), dp1 = d | 1, ldl = npmax | 1, dp = d + 1;
    double* Xp = sm;
    double* Lm = Xp + npmax * dp1;
    double* rhs = Lm + npmax * ldl;
    double* sv = rhs + npmax;
    int np = st->np[a];
    __syncwarp();
    for (int inner = 0; inner < 3 * kAaMaxP && np > 0; ++inner) {
        for (int e = lane; e < np * d; e += 32) {
            const int p = e / d, t = e - p * d;
            Xp[p * dp1 + t] = X[(int64_t)P[p] * d + t];
        }
        __syncwarp();
        for (int p = 0; p < np; ++p)
            for (int q = lane; q <= p; q += 32) {
                double v = M * M;
                for (int t = 0; t < d; ++t) v += Xp[p * dp1 + t] * Xp[q * dp1 + t];
                Lm[p * ldl + q] = v;
            }
        for (int p = lane; p < np; p += 32) {
            double v = M * M;
            for (int t = 0; t < d; ++t) v += Xp[p * dp1 + t] * Zhat[a * d + t];
            rhs[p] = v;
        }
        __syncwarp();
        bool ok = true;
        for (int j = 0; j < np; ++j) {                  // Cholesky in place, column j: diagonal, then the rows below it
            double v = Lm[j * ldl + j];
            for (int k = 0; k < j; ++k) v -= Lm[j * ldl + k] * Lm[j * ldl + k];
            if (!(v > 0.0)) { ok = false; break; }      // the same value in every lane
            const double ljj = sqrt(v);
            __syncwarp();
            if (lane == 0) Lm[j * ldl + j] = ljj;
            for (int i = j + 1 + lane; i < np; i += 32) {
                double w = Lm[i * ldl + j];
                for (int k = 0; k < j; ++k) w -= Lm[i * ldl + k] * Lm[j * ldl + k];
                Lm[i * ldl + j] = w / ljj;
            }
            __syncwarp();
        }
        if (!ok) {                                      // numerically dependent point: drop the newest, stop adding
            --np;
            if (lane == 0) st->done[a] = 1;
            __syncwarp();
            continue;
        }
        for (int i = lane; i < np; i += 32) sv[i] = rhs[i];
        __syncwarp();
        for (int k = 0; k < np; ++k) {                  // L y = rhs
            const double sk = sv[k] / Lm[k * ldl + k];
            __syncwarp();
            if (lane == 0) sv[k] = sk;
            for (int i = k + 1 + lane; i < np; i += 32) sv[i] -= Lm[i * ldl + k] * sk;
            __syncwarp();
        }
        for (int k = np - 1; k >= 0; --k) {             // L^T s = y
            const double sk = sv[k] / Lm[k * ldl + k];
            __syncwarp();
            if (lane == 0) sv[k] = sk;
            for (int i = lane; i < k; i += 32) sv[i] -= Lm[k * ldl + i] * sk;
            __syncwarp();
        }
        double smin = 1e300;
        for (int p = lane; p < np; p += 32) smin = fmin(smin, sv[p]);
        for (int o = 16; o > 0; o >>= 1) smin = fmin(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        if (smin > 0.0) {
            for (int p = lane; p < np; p += 32) b[p] = sv[p];
            __syncwarp();
            break;
        }
        double alpha = 1e300;
        for (int p = lane; p < np; p += 32)
            if (sv[p] <= 0.0) alpha = fmin(alpha, b[p] / (b[p] - sv[p]));
        for (int o = 16; o > 0; o >>= 1) alpha = fmin(alpha, __shfl_xor_sync(0xffffffffu, alpha, o));
        for (int p = lane; p < np; p += 32) b[p] += alpha * (sv[p] - b[p]);
        __syncwarp();
        int q = 0;
        if (lane == 0) {
            for (int p = 0; p < np; ++p) {
                if (b[p] <= 1e-15 && sv[p] <= 0.0) continue;
                P[q] = P[p]; b[q] = b[p]; ++q;
            }
        }
        np = __shfl_sync(0xffffffffu, q, 0);
        __syncwarp();
    }
    if (lane == 0) st->np[a] = np;
    for (int t = lane; t < dp; t += 32) {
        double v = t < d ? Zhat[a * d + t] : M;
        for (int p = 0; p < np; ++p) v -= b[p] * (t < d ? X[(int64_t)P[p] * d + t] : M);
        R[a * dp + t] = v;
    }
    __syncwarp();
}

// warm start of the beta step: every archetype (one warp each) keeps the passive points of the previous iteration
__global__ void __launch_bounds__(32)
aa_beta_refit_kernel(const double* __restrict__ X, int d, double M, const double* __restrict__ Zhat, int na,
                     BetaState* __restrict__ st, int* __restrict__ pidx, double* __restrict__ bcoef, double* __restrict__ R) {
    extern __shared__ double sh[];
    const int a = blockIdx.x;
    if (a >= na) return;
    beta_refit_warp(X, d, M, Zhat, a, st, pidx + a * kAaMaxP, bcoef + a * kAaMaxP, R, sh);
}

// one warp (= one CTA) per archetype: final arg-max, add to the passive set, Lawson-Hanson inner loop, new residual
__global__ void __launch_bounds__(32)
aa_beta_update_kernel(const double* __restrict__ X, int64_t n, int d, double M, const double* __restrict__ Zhat, int na,
                      int nblocks, const double* __restrict__ cand_w, const int64_t* __restrict__ cand_i,
                      BetaState* __restrict__ st, int* __restrict__ pidx, double* __restrict__ bcoef, double* __restrict__ R) {
    extern __shared__ double sh[];
    const int a = blockIdx.x, lane = threadIdx.x & 31;
    if (a >= na || st->all_done) return;
    if (!st->done[a]) {
        // final arg-max over the CTAs of the gradient pass (lane-strided, lowest index wins ties)
        double bw = -1e300;
        int64_t bi = -1;
        for (int b = lane; b < nblocks; b += 32) {
            const double w = cand_w[(size_t)b * na + a];
            const int64_t i = cand_i[(size_t)b * na + a];
            if (w > bw || (w == bw && i >= 0 && (bi < 0 || i < bi))) { bw = w; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ow = __shfl_xor_sync(0xffffffffu, bw, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ow > bw || (ow == bw && oi >= 0 && (bi < 0 || oi < bi))) { bw = ow; bi = oi; }
        }
        int* P = pidx + a * kAaMaxP;
        double* b = bcoef + a * kAaMaxP;
        int add = 0;
        if (lane == 0) {
            const int np = st->np[a];
            // scale for the stopping test: |[zhat, M]| * max |[x, M]| ~ M^2
            double znorm = M * M;
            for (int t = 0; t < d; ++t) znorm += Zhat[a * d + t] * Zhat[a * d + t];
            const double tol = 1e-11 * znorm;
            bool in_set = false;
            for (int p = 0; p < np; ++p) in_set |= (P[p] == (int)bi);
            if (bi < 0 || bw <= tol || in_set || np >= beta_npmax(d) - 1) {
                st->done[a] = 1;
            } else {
                P[np] = (int)bi;
                b[np] = 0.0;
                st->np[a] = np + 1;
                add = 1;
            }
        }
        add = __shfl_sync(0xffffffffu, add, 0);
        if (add) beta_refit_warp(X, d, M, Zhat, a, st, P, b, R, sh);
    }
    (void)n;
}

__global__ void aa_beta_check_kernel(BetaState* st, int na) {
    if (threadIdx.x == 0) {
        int all = 1;
        for (int a = 0; a < na; ++a) all &= st->done[a];
        st->all_done = all;
        st->iters += 1;
    }
}

__global__ void aa_beta_init_kernel(BetaState* st, const double* __restrict__ Zhat, int na, int d, double M, int warm,
                                    double* __restrict__ R) {
    const int t = threadIdx.x;
    if (t == 0) { st->all_done = 0; st->iters = 0; }
    if (t < na) { if (!warm) st->np[t] = 0; st->done[t] = 0; }
    for (int e = t; e < na * (d + 1); e += blockDim.x) {
        const int a = e / (d + 1), q = e % (d + 1);
        R[e] = q < d ? Zhat[a * d + q] : M;
    }
}

// Z[a] = sum_p coef_p x_p with coef = max(b,0) / sum;  also writes the sparse betas (index, coef) lists
__global__ void aa_beta_finish_kernel(const double* __restrict__ X, int d, int na, const BetaState* __restrict__ st,
                                      const int* __restrict__ pidx, double* __restrict__ bcoef, double* __restrict__ Z) {
    const int a = blockIdx.x;
    if (a >= na) return;
    const int np = st->np[a];
    const int* P = pidx + a * kAaMaxP;
    double* b = bcoef + a * kAaMaxP;
    __shared__ double inv;
    if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int p = 0; p < np; ++p) { b[p] = fmax(b[p], 0.0); sum += b[p]; }
        inv = 1.0 / sum;
        for (int p = 0; p < np; ++p) b[p] *= inv;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < d; t += blockDim.x) {
        double v = 0.0;
        for (int p = 0; p < np; ++p) v += b[p] * X[(int64_t)P[p] * d + t];
        Z[a * d + t] = v;
    }
}

// rss partials, rows read straight from global memory (fallback when the staged rows do not fit shared memory)
__global__ void __launch_bounds__(256)
aa_rss_direct_kernel(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ alphas, const double* __restrict__ Z,
                     int na, double* __restrict__ part) {
    extern __shared__ double sh[];
    double* Zs = sh;
    __shared__ double red[32];
    for (int e = threadIdx.x; e < na * d; e += blockDim.x) Zs[e] = Z[e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < n) {
        const double* __restrict__ x = X + i * d;
        const double* __restrict__ al = alphas + i * na;
        for (int t = 0; t < d; ++t) {
            double v = x[t];
            for (int a = 0; a < na; ++a) v -= al[a] * Zs[a * d + t];
            r += v * v;
        }
    }
    r = warp_sum(r);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}

// rss partials for n_archetypes <= 16: the spot's coefficients stay in registers, Z is kept transposed so that the
// archetypes of one coordinate come with 16-byte loads (same per-spot arithmetic as aa_rss_kernel)
template <int NA>
__global__ void __launch_bounds__(256)
aa_rss_fixed_kernel(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ alphas, const double* __restrict__ Z,
                    int na, double* __restrict__ part) {
    extern __shared__ __align__(16) double sh[];
    const int dpad = d | 1, apad = na | 1;
    double* Zt = sh;                       // d x NA, zero on the padding
    double* Xs = Zt + NA * d;              // 256 x dpad
    double* As = Xs + 256 * dpad;          // 256 x apad
    __shared__ double red[32];
    const int64_t i_base = (int64_t)blockIdx.x * 256;
    const int nvalid = (int)((n - i_base) < 256 ? (n - i_base) : 256);
    for (int e = threadIdx.x; e < NA * d; e += 256) {
        const int t = e / NA, a = e - t * NA;
        Zt[e] = a < na ? Z[a * d + t] : 0.0;
    }
    stage_rows(Xs, dpad, X + i_base * d, d, nvalid);
    stage_rows(As, apad, alphas + i_base * na, na, nvalid);
    stage_wait();
    __syncthreads();
    double r = 0.0;
    if ((int)threadIdx.x < nvalid) {
        const double* __restrict__ x = Xs + threadIdx.x * dpad;
        double al[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) al[a] = a < na ? As[threadIdx.x * apad + a] : 0.0;
        for (int t = 0; t < d; ++t) {
            double v = x[t];
            const double2* __restrict__ z2 = reinterpret_cast<const double2*>(Zt + t * NA);
#pragma unroll
            for (int h = 0; h < NA / 2; ++h) {
                const double2 z = z2[h];
                v -= al[2 * h] * z.x;
                v -= al[2 * h + 1] * z.y;
            }
            r += v * v;
        }
    }
    r = warp_sum(r);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}

// rss partials:  sum_i || x_i - sum_a alpha_ia Z_a ||^2.  The CTA's 256 rows of X and of alphas are contiguous blocks:
// they are staged in shared memory with coalesced loads (odd row strides: conflict-free per-thread rows)
__global__ void __launch_bounds__(256)
aa_rss_kernel(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ alphas, const double* __restrict__ Z,
              int na, double* __restrict__ part) {
    extern __shared__ double sh[];
    const int dpad = d | 1, apad = na | 1;
    double* Zs = sh;                       // na x d
    double* Xs = Zs + na * d;              // 256 x dpad
    double* As = Xs + 256 * dpad;          // 256 x apad
    __shared__ double red[32];
    const int64_t i_base = (int64_t)blockIdx.x * 256;
    const int nvalid = (int)((n - i_base) < 256 ? (n - i_base) : 256);
    for (int e = threadIdx.x; e < na * d; e += 256) Zs[e] = Z[e];
    stage_rows(Xs, dpad, X + i_base * d, d, nvalid);
    stage_rows(As, apad, alphas + i_base * na, na, nvalid);
    stage_wait();
    __syncthreads();
    double r = 0.0;
    if ((int)threadIdx.x < nvalid) {
        const double* __restrict__ x = Xs + threadIdx.x * dpad;
        const double* __restrict__ al = As + threadIdx.x * apad;
        for (int t = 0; t < d; ++t) {
            double v = x[t];
            for (int a = 0; a < na; ++a) v -= al[a] * Zs[a * d + t];
            r += v * v;
        }
    }
    r = warp_sum(r);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}

// ||X||_F^2 partials (once per restart): 256 rows per CTA, fixed order
__global__ void __launch_bounds__(256) aa_xnorm_kernel(const double* __restrict__ X, int64_t n, int d, double* __restrict__ part) {
    __shared__ double red[8];
    const int64_t e0 = (int64_t)blockIdx.x * 256 * d, e1 = e0 + (int64_t)256 * d < n * d ? e0 + (int64_t)256 * d : n * d;
    double t = 0.0;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += 256) t += X[e] * X[e];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double u = 0.0;
        for (int w = 0; w < 8; ++w) u += red[w];
        part[blockIdx.x] = u;
    }
}
// rss = ||X - A Z||_F^2 = ||X||^2 - 2 <A^T X, Z> + <A^T A, Z Z^T>  from the Gram sums the alpha step leaves behind (one CTA, fixed
// order).  The three terms are 10 - 100 times the result: absolute error ~1e-15 ||X||^2, five orders below the 1e-3 stop rule.
__global__ void __launch_bounds__(256)
aa_rss_gram_kernel(const double* __restrict__ gram, const double* __restrict__ Z, int na, int d, const double* __restrict__ xnorm2,
                   double* __restrict__ rss_out) {
    __shared__ double red[256];
    const double* __restrict__ S = gram;                 // A^T A  [na][na]
    const double* __restrict__ B = gram + na * na;       // A^T X  [na][d]
    double t = 0.0;
    for (int e = threadIdx.x; e < na * d; e += 256) t -= 2.0 * B[e] * Z[e];
    for (int e = threadIdx.x; e < na * na; e += 256) {
        const int a = e / na, b = e - a * na;
        double zz = 0.0;
        for (int q = 0; q < d; ++q) zz += Z[a * d + q] * Z[b * d + q];
        t += S[e] * zz;
    }
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *rss_out = *xnorm2 + red[0];
}

__global__ void __launch_bounds__(256) aa_sum_kernel(const double* __restrict__ part, int np, double* __restrict__ out) {
    // fixed order: every thread sums a contiguous chunk, thread 0 adds the 256 chunk sums in order
    __shared__ double red[256];
    const int chunk = (np + 255) / 256;
    const int p0 = threadIdx.x * chunk, p1 = p0 + chunk < np ? p0 + chunk : np;
    double t = 0.0;
    for (int p = p0; p < p1; ++p) t += part[p];
    red[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x < 32) {
        double u = 0.0;
        for (int w = 0; w < 8; ++w) u += red[threadIdx.x * 8 + w];
        // lanes hold consecutive groups of 8 chunks: a fixed binary tree over the 32 lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
        if (threadIdx.x == 0) *out = u;
    }
}

// a row index outside [0, n) never touches X: the archetype becomes NaN, which every later step carries to the result
__global__ void aa_gather_rows_kernel(const double* __restrict__ X, int64_t n, int d, const int64_t* __restrict__ rows, int na,
                                      double* __restrict__ Z) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= na * d) return;
    const int64_t r = rows[e / d];
    Z[e] = (r >= 0 && r < n) ? X[r * d + e % d] : __longlong_as_double(0x7ff8000000000000LL);
}

// ---- FurthestSum seeding (Morup & Hansen 2012; the initialisation archetypes 0.4.x draws) -----------------
//   total[i] += sign * || x_i - x_c ||   for one chosen point c; then arg-max over the points not chosen yet
__global__ void __launch_bounds__(256)
aa_fs_update_kernel(const double* __restrict__ X, int64_t n, int d, int64_t c, double sign, double* __restrict__ total) {
    extern __shared__ double xc[];
    for (int t = threadIdx.x; t < d; t += blockDim.x) xc[t] = X[c * d + t];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // same arithmetic as the host formulation: sqrt(max(|x_i|^2 - 2 x_i.x_c + |x_c|^2, 0))
    double sq = 0.0, dot = 0.0, sc = 0.0;
    const double* __restrict__ x = X + i * d;
    for (int t = 0; t < d; ++t) { sq += x[t] * x[t]; dot += x[t] * xc[t]; sc += xc[t] * xc[t]; }
    const double v = sq - 2.0 * dot + sc;
    total[i] += sign * sqrt(v > 0.0 ? v : 0.0);
}

// arg-max of total[] over the points with avail[i] != 0; lowest index wins ties (numpy.argmax semantics)
__global__ void __launch_bounds__(256)
aa_fs_argmax_partial_kernel(const double* __restrict__ total, const unsigned char* __restrict__ avail, int64_t n,
                            double* __restrict__ cand_w, int64_t* __restrict__ cand_i) {
    __shared__ double rw[8];
    __shared__ int64_t ri[8];
    double bw = -1e300;
    int64_t bi = -1;
    const int64_t i0 = (int64_t)blockIdx.x * 4096, i1 = i0 + 4096 < n ? i0 + 4096 : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x)
        if (avail[i] && (total[i] > bw || bi < 0)) { bw = total[i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
        const double ow = __shfl_xor_sync(0xffffffffu, bw, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ow > bw || (ow == bw && oi < bi))) { bw = ow; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { rw[warp] = bw; ri[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (ri[w] >= 0 && (bi < 0 || rw[w] > bw || (rw[w] == bw && ri[w] < bi))) { bw = rw[w]; bi = ri[w]; }
        cand_w[blockIdx.x] = bw;
        cand_i[blockIdx.x] = bi;
    }
}

__global__ void aa_fs_argmax_final_kernel(const double* __restrict__ cand_w, const int64_t* __restrict__ cand_i, int nb,
                                          unsigned char* __restrict__ avail, int64_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double bw = -1e300;
    int64_t bi = -1;
    for (int b = 0; b < nb; ++b)
        if (cand_i[b] >= 0 && (bi < 0 || cand_w[b] > bw || (cand_w[b] == bw && cand_i[b] < bi))) { bw = cand_w[b]; bi = cand_i[b]; }
    *out = bi;
    if (bi >= 0) avail[bi] = 0;
}

struct AaPlan {
    int alpha_blocks, grad_blocks, rss_blocks, nacc;
    size_t off_part, off_part2, off_zhat, off_R, off_state, off_pidx, off_bcoef, off_candw, off_candi, off_scr, off_rsspart, total;
    size_t off_gram, off_xnorm;     // reduced A^T A | A^T X of the iteration, ||X||_F^2 of the restart
};

static AaPlan aa_plan(int64_t n, int d, int na) {
    AaPlan p;
    p.alpha_blocks = (int)ceil_div<int64_t>(n, 128);
    p.grad_blocks = (int)ceil_div<int64_t>(n, 1024);
    p.rss_blocks = (int)ceil_div<int64_t>(n, 256);
    p.nacc = na * na + na * d;
    size_t o = 0;
    auto take = [&](size_t b) { size_t x = o; o += round_up<size_t>(b, 256); return x; };
    p.off_part = take((size_t)p.alpha_blocks * p.nacc * 8);
    p.off_part2 = take((size_t)ceil_div(p.alpha_blocks, kAaRedChunk) * p.nacc * 8);
    p.off_zhat = take((size_t)na * d * 8);
    p.off_R = take((size_t)na * (d + 1) * 8);
    p.off_state = take(sizeof(BetaState));
    p.off_pidx = take((size_t)na * kAaMaxP * 4);
    p.off_bcoef = take((size_t)na * kAaMaxP * 8);
    int64_t cand_rows = ceil_div<int64_t>(n, 128);          // staged gradient pass: one row per persistent CTA
    if (cand_rows > kNumSMs * 4) cand_rows = kNumSMs * 4;
    if (cand_rows < p.grad_blocks) cand_rows = p.grad_blocks;
    p.off_candw = take((size_t)cand_rows * na * 8);
    p.off_candi = take((size_t)cand_rows * na * 8);
    p.off_scr = take(256);                  // (unused: the beta refits keep their scratch in shared memory)
    p.off_rsspart = take((size_t)p.rss_blocks * 8);
    p.off_gram = take((size_t)p.nacc * 8);
    p.off_xnorm = take(8);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_aa_workspace_bytes(int64_t n, int d, int n_archetypes) {
    if (n <= 0 || d <= 0 || n_archetypes <= 0) return 1024;
    return chr::aa_plan(n, d, n_archetypes).total + 1024;
}

int chr_aa_init_archetypes(const double* X, int64_t n, int d, const int64_t* rows, int n_archetypes, double* Z,
                           chr_stream_t stream) {
    using namespace chr;
    CHR_REQUIRE(X && rows && Z && n > 0 && d > 0 && n_archetypes > 0, "chr_aa_init_archetypes: bad arguments");
    CHR_REQUIRE(d <= kAaMaxD && n_archetypes <= kAaMaxA, "chr_aa_init_archetypes: need d <= %d and n_archetypes <= %d (got %d, %d)",
                kAaMaxD, kAaMaxA, d, n_archetypes);
    aa_gather_rows_kernel<<<(unsigned)ceil_div(n_archetypes * d, 256), 256, 0, (cudaStream_t)stream>>>(X, n, d, rows, n_archetypes, Z);
    CHR_LAUNCH_CHECK();
    return 0;
}

size_t chr_aa_furthest_sum_workspace_bytes(int64_t n) {
    const int64_t nb = chr::ceil_div<int64_t>(n > 0 ? n : 1, 4096);
    return chr::round_up<size_t>((size_t)(n > 0 ? n : 1) * 8, 256) + chr::round_up<size_t>((size_t)(n > 0 ? n : 1), 256) +
           2 * chr::round_up<size_t>((size_t)nb * 8, 256) + 256 + 512;
}

// FurthestSum: `first` is the random starting point (drawn by the host RNG); rows_out (HOST, int64[n_archetypes])
// receives the chosen data points in the order the host algorithm produces them.
int chr_aa_furthest_sum(const double* X, int64_t n, int d, int n_archetypes, int64_t first, int64_t* rows_out,
                        void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(X && rows_out && workspace, "chr_aa_furthest_sum: null pointer argument");
    CHR_REQUIRE(n > 0 && d > 0 && n_archetypes >= 1 && n_archetypes < n && first >= 0 && first < n, "chr_aa_furthest_sum: bad arguments");
    CHR_REQUIRE(d <= kAaMaxD && n_archetypes <= kAaMaxA, "chr_aa_furthest_sum: need d <= %d and n_archetypes <= %d (got %d, %d)", kAaMaxD,
                kAaMaxA, d, n_archetypes);
    CHR_REQUIRE(workspace_bytes >= chr_aa_furthest_sum_workspace_bytes(n) - 256, "chr_aa_furthest_sum: workspace too small");
    const int nb = (int)ceil_div<int64_t>(n, 4096);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* total = (double*)ws;
    unsigned char* avail = (unsigned char*)(ws + round_up<size_t>((size_t)n * 8, 256));
    double* cand_w = (double*)((char*)avail + round_up<size_t>((size_t)n, 256));
    int64_t* cand_i = (int64_t*)((char*)cand_w + round_up<size_t>((size_t)nb * 8, 256));
    int64_t* pick = (int64_t*)((char*)cand_i + round_up<size_t>((size_t)nb * 8, 256));
    CHR_CUDA(cudaMemsetAsync(total, 0, (size_t)n * 8, st));
    CHR_CUDA(cudaMemsetAsync(avail, 1, (size_t)n, st));
    CHR_CUDA(cudaMemsetAsync(avail + first, 0, 1, st));
    const unsigned gu = (unsigned)ceil_div<int64_t>(n, 256);
    const size_t sm = (size_t)d * 8;
    // the host algorithm (chrysalis_b200/core.py:_furthest_sum): picks = [first]; for step in 1 .. A + 10:
    //   if step > A - 1: drop the oldest pick (subtract its distances, make it available again)
    //   add the distances to the current point; next = arg-max over available points; append
    int64_t picks[kAaMaxA + 16];                    // n_archetypes + 11 <= kAaMaxA + 11 entries (checked above)
    int np = 0, head = 0;
    picks[np++] = first;
    int64_t cur = first;
    for (int step = 1; step < n_archetypes + 11; ++step) {
        if (step > n_archetypes - 1) {
            const int64_t oldest = picks[head++];
            aa_fs_update_kernel<<<gu, 256, sm, st>>>(X, n, d, oldest, -1.0, total);
            CHR_CUDA(cudaMemsetAsync(avail + oldest, 1, 1, st));
        }
        aa_fs_update_kernel<<<gu, 256, sm, st>>>(X, n, d, cur, 1.0, total);
        aa_fs_argmax_partial_kernel<<<nb, 256, 0, st>>>(total, avail, n, cand_w, cand_i);
        aa_fs_argmax_final_kernel<<<1, 32, 0, st>>>(cand_w, cand_i, nb, avail, pick);
        CHR_LAUNCH_CHECK();
        CHR_CUDA(cudaMemcpyAsync(&cur, pick, 8, cudaMemcpyDeviceToHost, st));
        CHR_CUDA(cudaStreamSynchronize(st));
        CHR_REQUIRE(cur >= 0, "chr_aa_furthest_sum: no available point left");
        picks[np++] = cur;
    }
    for (int a = 0; a < n_archetypes; ++a) rows_out[a] = picks[head + a];
    return 0;
}

// one full iteration:  alphas <- NNLS(X | Z);  Z <- pinv(alphas) X;  betas <- NNLS(Z | X);  Z <- betas X;  rss
int chr_aa_iterate(const double* X, int64_t n, int d, int n_archetypes, double penalty_M, int warm, double* Z, double* alphas,
                   double* rss_out, int* beta_passes_out, void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    const int na = n_archetypes;
    CHR_REQUIRE(X && Z && alphas && rss_out && workspace, "chr_aa_iterate: null pointer argument");
    CHR_REQUIRE(n > 0 && d > 0 && d <= kAaMaxD && na >= 2 && na <= kAaMaxA, "chr_aa_iterate: need 0 < d <= %d, 2 <= n_archetypes <= %d",
                kAaMaxD, kAaMaxA);
    AaPlan p = aa_plan(n, d, na);
    CHR_REQUIRE(workspace_bytes >= p.total + 256, "chr_aa_iterate: workspace too small (%zu < %zu)", workspace_bytes, p.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* part = (double*)(ws + p.off_part);
    double* part2 = (double*)(ws + p.off_part2);
    double* Zhat = (double*)(ws + p.off_zhat);
    double* R = (double*)(ws + p.off_R);
    BetaState* bst = (BetaState*)(ws + p.off_state);
    int* pidx = (int*)(ws + p.off_pidx);
    double* bcoef = (double*)(ws + p.off_bcoef);
    double* candw = (double*)(ws + p.off_candw);
    int64_t* candi = (int64_t*)(ws + p.off_candi);
    double* rsspart = (double*)(ws + p.off_rsspart);

    timer_begin(CHR_FAMILY_AA, timing_env(), st);
    // ---- alphas
    const size_t sm_alpha = (size_t)(na * na + na * d + p.nacc + 128 * na) * 8;
    if (na <= 16 && !getenv("CHR_AA_ALPHA_GENERAL")) {
        const int NA = (na + 3) & ~3;
        const size_t region = aa_alpha_fixed_region_doubles(NA, na, d);
        const size_t smf = ((size_t)NA * NA + (size_t)NA * d + region + (size_t)128 * (na | 1)) * 8;
#define CHR_ALPHA_FIXED(NAV)                                                                                              \
    do {                                                                                                                  \
        CHR_CUDA(cudaFuncSetAttribute(aa_alpha_fixed_kernel<NAV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf)); \
        aa_alpha_fixed_kernel<NAV><<<p.alpha_blocks, 128, smf, st>>>(X, n, d, Z, na, penalty_M, warm, alphas, part, (int)region); \
    } while (0)
        if (NA == 4) CHR_ALPHA_FIXED(4);
        else if (NA == 8) CHR_ALPHA_FIXED(8);
        else if (NA == 12) CHR_ALPHA_FIXED(12);
        else CHR_ALPHA_FIXED(16);
#undef CHR_ALPHA_FIXED
    } else if (na <= 16) aa_alpha_kernel<16><<<p.alpha_blocks, 128, sm_alpha, st>>>(X, n, d, Z, na, penalty_M, warm, alphas, part);
    else if (na <= 32) {
        CHR_CUDA(cudaFuncSetAttribute(aa_alpha_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_alpha));
        aa_alpha_kernel<32><<<p.alpha_blocks, 128, sm_alpha, st>>>(X, n, d, Z, na, penalty_M, warm, alphas, part);
    } else {
        CHR_CUDA(cudaFuncSetAttribute(aa_alpha_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_alpha));
        aa_alpha_kernel<64><<<p.alpha_blocks, 128, sm_alpha, st>>>(X, n, d, Z, na, penalty_M, warm, alphas, part);
    }
    CHR_LAUNCH_CHECK();
    // ---- Zhat = pinv(alphas) X
    const size_t sm_pinv = (size_t)(2 * na * na + na * d + 2 * na + 8) * 8;      // + rotation table of a Jacobi round
    if (sm_pinv > 48 * 1024) CHR_CUDA(cudaFuncSetAttribute(aa_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pinv));
    const int nred = ceil_div(p.alpha_blocks, kAaRedChunk);
    aa_partial_reduce_kernel<<<dim3((unsigned)ceil_div(p.nacc, 128), (unsigned)nred), 128, 0, st>>>(part, p.alpha_blocks, p.nacc, part2);
    double* gram = (double*)(ws + p.off_gram);
    double* xnorm2 = (double*)(ws + p.off_xnorm);
    aa_pinv_kernel<<<1, 256, sm_pinv, st>>>(part2, nred, na, d, n, Zhat, gram);
    CHR_LAUNCH_CHECK();
    // ---- betas: active-set passes, every archetype advances one step per pass
    aa_beta_init_kernel<<<1, 256, 0, st>>>(bst, Zhat, na, d, penalty_M, warm, R);
    const size_t sm_refit = beta_refit_smem_bytes(d);
    if (sm_refit > 48 * 1024) {
        CHR_CUDA(cudaFuncSetAttribute(aa_beta_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_refit));
        CHR_CUDA(cudaFuncSetAttribute(aa_beta_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_refit));
    }
    if (warm) aa_beta_refit_kernel<<<na, 32, sm_refit, st>>>(X, d, penalty_M, Zhat, na, bst, pidx, bcoef, R);
    const size_t sm_grad = (size_t)na * (d + 1) * 8 + 32 * 8 + 32 * 8;
    const size_t sm_gradf = ((size_t)(d + 1) * 16 + kGradChunk * (size_t)(d | 1) + 4 * 16 * 2) * 8;
    const int NAg = (na + 3) & ~3;
    if (na <= 16 && sm_gradf > 48 * 1024) {
        if (NAg == 4) CHR_CUDA(cudaFuncSetAttribute(aa_beta_grad_fixed_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gradf));
        else if (NAg == 8) CHR_CUDA(cudaFuncSetAttribute(aa_beta_grad_fixed_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gradf));
        else if (NAg == 12) CHR_CUDA(cudaFuncSetAttribute(aa_beta_grad_fixed_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gradf));
        else CHR_CUDA(cudaFuncSetAttribute(aa_beta_grad_fixed_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gradf));
    }
    // persistent grid of the staged gradient pass (aa_plan sizes the candidate arrays for it)
    const int64_t grad_chunks = ceil_div<int64_t>(n, kGradChunk);
    int gradf_grid = kNumSMs * 3;
    if (gradf_grid > grad_chunks) gradf_grid = (int)grad_chunks;
    const int cand_blocks = na <= 16 ? gradf_grid : p.grad_blocks;
    const int max_pass = 3 * (d + 2) + 8;
    int passes = 0;
    int host_done = 0;
    while (passes < max_pass && !host_done) {
        const int batch = warm ? 2 : (passes == 0 ? (d + 2 < 8 ? d + 2 : 8) : 4);   // cold: first look after a handful of passes
        for (int b = 0; b < batch && passes < max_pass; ++b, ++passes) {
            if (NAg == 4) aa_beta_grad_fixed_kernel<4><<<gradf_grid, kGradThreads, sm_gradf, st>>>(X, n, d, penalty_M, R, na, bst, candw, candi);
            else if (NAg == 8) aa_beta_grad_fixed_kernel<8><<<gradf_grid, kGradThreads, sm_gradf, st>>>(X, n, d, penalty_M, R, na, bst, candw, candi);
            else if (NAg == 12) aa_beta_grad_fixed_kernel<12><<<gradf_grid, kGradThreads, sm_gradf, st>>>(X, n, d, penalty_M, R, na, bst, candw, candi);
            else if (NAg == 16) aa_beta_grad_fixed_kernel<16><<<gradf_grid, kGradThreads, sm_gradf, st>>>(X, n, d, penalty_M, R, na, bst, candw, candi);
            else
                aa_beta_grad_kernel<<<p.grad_blocks, 256, sm_grad, st>>>(X, n, d, penalty_M, R, na, bst, pidx, candw, candi);
            aa_beta_update_kernel<<<na, 32, sm_refit, st>>>(X, n, d, penalty_M, Zhat, na, cand_blocks, candw, candi, bst, pidx, bcoef, R);
            aa_beta_check_kernel<<<1, 32, 0, st>>>(bst, na);
        }
        CHR_LAUNCH_CHECK();
        CHR_CUDA(cudaMemcpyAsync(&host_done, &bst->all_done, sizeof(int), cudaMemcpyDeviceToHost, st));
        CHR_CUDA(cudaStreamSynchronize(st));
    }
    if (beta_passes_out) *beta_passes_out = passes;
    aa_beta_finish_kernel<<<na, 64, 0, st>>>(X, d, na, bst, pidx, bcoef, Z);
    // ---- rss: from the iteration's Gram sums (no pass over X); CHR_AA_RSS_DIRECT=1 keeps the per-spot residual pass
    if (!getenv("CHR_AA_RSS_DIRECT")) {
        if (!warm) {                       // first iteration of a restart: ||X||_F^2 into the workspace
            aa_xnorm_kernel<<<p.rss_blocks, 256, 0, st>>>(X, n, d, rsspart);
            aa_sum_kernel<<<1, 256, 0, st>>>(rsspart, p.rss_blocks, xnorm2);
        }
        aa_rss_gram_kernel<<<1, 256, 0, st>>>(gram, Z, na, d, xnorm2, rss_out);
        timer_end(CHR_FAMILY_AA, timing_env(), st);
        CHR_LAUNCH_CHECK();
        return 0;
    }
    const size_t sm_rss = ((size_t)na * d + 256 * (size_t)(d | 1) + 256 * (size_t)(na | 1)) * 8;
    const int NAr = (na + 3) & ~3;
    const size_t sm_rssf = ((size_t)NAr * d + 256 * (size_t)(d | 1) + 256 * (size_t)(na | 1)) * 8;
    if (na <= 16 && sm_rssf <= 160 * 1024) {
#define CHR_RSS_FIXED(NAV)                                                                                                  \
    do {                                                                                                                    \
        if (sm_rssf > 48 * 1024)                                                                                            \
            CHR_CUDA(cudaFuncSetAttribute(aa_rss_fixed_kernel<NAV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_rssf)); \
        aa_rss_fixed_kernel<NAV><<<p.rss_blocks, 256, sm_rssf, st>>>(X, n, d, alphas, Z, na, rsspart);                       \
    } while (0)
        if (NAr == 4) CHR_RSS_FIXED(4);
        else if (NAr == 8) CHR_RSS_FIXED(8);
        else if (NAr == 12) CHR_RSS_FIXED(12);
        else CHR_RSS_FIXED(16);
#undef CHR_RSS_FIXED
    } else if (sm_rss <= 160 * 1024) {
        if (sm_rss > 48 * 1024) CHR_CUDA(cudaFuncSetAttribute(aa_rss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_rss));
        aa_rss_kernel<<<p.rss_blocks, 256, sm_rss, st>>>(X, n, d, alphas, Z, na, rsspart);
    } else {
        aa_rss_direct_kernel<<<p.rss_blocks, 256, (size_t)na * d * 8, st>>>(X, n, d, alphas, Z, na, rsspart);
    }
    aa_sum_kernel<<<1, 256, 0, st>>>(rsspart, p.rss_blocks, rss_out);
    timer_end(CHR_FAMILY_AA, timing_env(), st);
    CHR_LAUNCH_CHECK();
    return 0;
}

#ifdef CHR_AA_STATS
// measurement aid of -DCHR_AA_STATS builds (ab_libs/), absent from the product library
int chr_debug_aa_stats(unsigned long long* out4, int reset) {
    using namespace chr;
    if (out4 && cudaMemcpyFromSymbol(out4, g_aa_stats, sizeof(unsigned long long) * 4) != cudaSuccess) return 2;
    if (reset) {
        const unsigned long long z[4] = {0, 0, 0, 0};
        if (cudaMemcpyToSymbol(g_aa_stats, z, sizeof(z)) != cudaSuccess) return 2;
    }
    return 0;
}
#endif

}  // extern "C"
