// Local Moran's I (LISA) with conditional permutation inference - SURVEY.md 8(f) rank 4.
//
// Replaces  esda.moran.Moran_Local(y, w, transformation='r', permutations=P)  as the authors call it
// next to the global statistic (/root/reference/article/A2_human_lymph_node/morans_i.ipynb:150);
// chrysalis/core.py itself never calls it.  For one variable y over the spots:
//     z = (y - mean) / std (population),   lag_i = (1/k) sum_{j in N(i)} z_j,
//     Is_i = (n - 1) z_i lag_i / sum z^2,  quadrant q_i in {1 HH, 2 LH, 3 LL, 4 HL}
// and, per spot, P conditional draws: the k neighbour values are replaced by k distinct values of the other
// n - 1 spots; p_sim_i = (min(#{sim >= Is_i}, P - #) + 1) / (P + 1).
// One thread per spot; draws come from a counter-based generator keyed by (seed, spot, draw, slot), so the
// result does not depend on the launch geometry.
#include "common.cuh"

namespace chr {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {   // splitmix64 finaliser, upper 32 bits
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}

__global__ void lisa_kernel(const float* __restrict__ z, int64_t n, const int32_t* __restrict__ nbr, int k, double z2ss,
                            float* __restrict__ lag_out, float* __restrict__ Is_out, int32_t* __restrict__ q_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += z[nbr[i * k + j]];
    const float lag = s / (float)k;
    lag_out[i] = lag;
    Is_out[i] = (float)((double)(n - 1) * (double)z[i] * (double)lag / z2ss);
    const bool zp = z[i] > 0.f, lp = lag > 0.f;
    q_out[i] = zp ? (lp ? 1 : 4) : (lp ? 2 : 3);
}

__global__ void lisa_perm_kernel(const float* __restrict__ z, int64_t n, int k, const float* __restrict__ Is, double z2ss, int P,
                                 uint64_t seed, float* __restrict__ p_sim, float* __restrict__ ei_sim, float* __restrict__ vi_sim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double scale = (double)(n - 1) * (double)z[i] / z2ss / (double)k;
    const float obs = Is[i];
    int larger = 0;
    double s1 = 0.0, s2 = 0.0;
    int32_t pick[64];
    const int kk = k < 64 ? k : 64;
    for (int p = 0; p < P; ++p) {
        float acc = 0.f;
        for (int j = 0; j < kk; ++j) {
            // j-th of k distinct other spots: redraw on the (rare) collision with an earlier pick
            int32_t c;
            for (uint32_t attempt = 0;; ++attempt) {
                const uint64_t key = seed ^ ((uint64_t)i * 0xD1B54A32D192ED03ull) ^ ((uint64_t)p << 40) ^ ((uint64_t)j << 24) ^ attempt;
                c = (int32_t)(((uint64_t)mix32(key) * (uint64_t)(n - 1)) >> 32);     // uniform in [0, n - 1)
                if (c >= i) ++c;                                                      // skip the spot itself
                bool dup = false;
                for (int e = 0; e < j; ++e) dup |= (pick[e] == c);
                if (!dup || attempt > 16) break;
            }
            pick[j] = c;
            acc += z[c];
        }
        const double sim = scale * (double)acc;
        larger += (sim >= (double)obs);
        s1 += sim;
        s2 += sim * sim;
    }
    if (P - larger < larger) larger = P - larger;
    p_sim[i] = (float)((larger + 1.0) / (P + 1.0));
    const double m = s1 / P;
    ei_sim[i] = (float)m;
    vi_sim[i] = (float)fmax(s2 / P - m * m, 0.0);
}

}  // namespace chr

extern "C" {

int chr_local_moran_f32(const float* z, int64_t n, const int32_t* nbr, int k, double z2ss, int permutations, uint64_t seed,
                        float* lag_out, float* Is_out, int32_t* q_out, float* p_sim_out, float* ei_sim_out, float* vi_sim_out,
                        chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(z && nbr && lag_out && Is_out && q_out, "chr_local_moran_f32: null pointer argument");
    CHR_REQUIRE(n > 2 && k > 0 && k < n && k <= 64, "chr_local_moran_f32: need n > 2 and 0 < k <= min(64, n - 1) (n=%lld, k=%d)", (long long)n, k);
    CHR_REQUIRE(z2ss > 0.0, "chr_local_moran_f32: the variable is constant");
    CHR_REQUIRE(permutations == 0 || (p_sim_out && ei_sim_out && vi_sim_out), "chr_local_moran_f32: permutation outputs missing");
    const unsigned grid = (unsigned)ceil_div<int64_t>(n, 128);
    lisa_kernel<<<grid, 128, 0, st>>>(z, n, nbr, k, z2ss, lag_out, Is_out, q_out);
    CHR_LAUNCH_CHECK();
    if (permutations > 0) {
        lisa_perm_kernel<<<grid, 128, 0, st>>>(z, n, k, Is_out, z2ss, permutations, seed, p_sim_out, ei_sim_out, vi_sim_out);
        CHR_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
