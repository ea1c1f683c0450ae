// K1b - the device representation of the row-standardised kNN weights W that the Moran kernel
// streams against ("GPU CSR build of W").  Input: nbr[n][k], the kNN lists that replace
//     w = weights.KNN.from_array(points, k=neighbors); w.transform = 'R'
// of /root/reference/chrysalis/core.py:64-65 (every weight is 1/k).  Built once per W, reused for
// every gene / gene shard / permutation.
//
// z'Az needs every UNORDERED pair {i, j} once:
//        z'Az = sum_{i<j} (a_ij + a_ji) z_i z_j + sum_i a_ii z_i^2
// so the adjacency counts are re-expressed over QUADS of 4 consecutive spots (consecutive spots
// of the Hilbert order are spatial neighbours):
//   * pairs inside a quad  -> 6 coefficients (the kernel holds the 4 rows in registers)
//   * every other pair     -> owned by exactly ONE endpoint, as (row, coefficient) in that spot's
//     list.  Pairs only one endpoint lists stay with it; mutual pairs start with a hash-chosen
//     owner and are moved by a few Jacobi sweeps so that the 4 lists of a quad get nearly equal
//     length R: the kernel walks a quad in R rounds of 4 rows.
//   * list entries are split into rows inside the quad's 112-spot tile (served from the tile the
//     kernel stages in shared memory) and rows outside it (global loads): R = R_loc + R_glob.
// On a square lattice with k = 8 a spot reads ~2.9 list rows instead of 8.  Any list is legal
// (duplicates, self loops, asymmetric kNN): coefficients are multiplicities.  k <= 64 (one
// 64-bit owner mask per spot).
#include <cstdlib>
#include "moran.cuh"

namespace chr {

constexpr int kRelaxIters = 8;        // sweeps that balance the per-spot list lengths
constexpr int kRelaxMaxIters = 40;    // further sweeps that lower the per-quad maxima (what the kernel's round count is)

__device__ __forceinline__ uint32_t pair_hash(uint32_t i, uint32_t j) {
    const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
    uint32_t h = (lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

// own[i] bit a: spot i owns the pair {i, nbr[i][a]}.  Only first occurrences of a neighbour carry a
// bit; quad-internal and self pairs carry none.  cnt[i] = list entries of spot i: in-tile | out-of-tile << 16
// (the kernel walks a quad in max(in-tile) + max(out-of-tile) rounds, so the two kinds are balanced separately).
__global__ void plan_init_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, uint64_t* __restrict__ own,
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
        if (j == (int32_t)i) { ++c; continue; }                        // self loop: one in-tile list entry, no bit
        if ((j >> 2) == (int32_t)(i >> 2)) continue;                   // quad-internal
        const int32_t* B = nbr + (int64_t)j * k;
        bool back = false;
        for (int b = 0; b < k; ++b) back |= (B[b] == (int32_t)i);
        bool mine = true;                                              // a pair only i lists stays with i
        if (back) {
            const uint32_t lo = (uint32_t)(i < j ? i : j), hi = (uint32_t)(i < j ? j : i);
            mine = ((pair_hash((uint32_t)i, (uint32_t)j) & 1u) ? lo : hi) == (uint32_t)i;
        }
        if (mine) { c += (j / kTileSpots) == (int32_t)(i / kTileSpots) ? 1 : (1 << 16); bits |= 1ull << a; }
    }
    own[i] = bits;
    cnt[i] = c;
}

// longest list (in-tile | out-of-tile << 16, each taken separately) among the spots of the quad of spot i
__device__ __forceinline__ int quad_max_counts(const int32_t* __restrict__ cnt, int64_t n, int64_t i) {
    int ml = 0, mg = 0;
    const int64_t q = i & ~(int64_t)3;
    for (int t = 0; t < 4; ++t)
        if (q + t < n) {
            const int c = cnt[q + t];
            ml = max(ml, c & 0xffff);
            mg = max(mg, (c >> 16) & 0xffff);
        }
    return ml | (mg << 16);
}
// longest combined list among the spots of the quad of spot i
__device__ __forceinline__ int quad_max_total(const int32_t* __restrict__ cnt, int64_t n, int64_t i) {
    int m = 0;
    const int64_t q = i & ~(int64_t)3;
    for (int t = 0; t < 4; ++t)
        if (q + t < n) m = max(m, (cnt[q + t] & 0xffff) + ((cnt[q + t] >> 16) & 0xffff));
    return m;
}

// one Jacobi sweep of owner balancing: a mutual pair moves to the other endpoint when that lowers
// the larger of the two list lengths; both endpoints evaluate the same rule on the same data.
// Sweeps >= kRelaxIters also move an entry one step downhill when its owner is the longest list of its
// quad and the receiver stays within its own quad's longest: the kernel pays max-over-the-quad rounds.
__global__ void plan_relax_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int iter, int combined, const uint64_t* __restrict__ own_in,
                                  const int32_t* __restrict__ cnt_in, uint64_t* __restrict__ own_out, int32_t* __restrict__ cnt_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* A = nbr + i * k;
    const uint64_t mine_old = own_in[i];
    uint64_t bits = mine_old;
    int c = cnt_in[i];
    const int ci = c;
    const bool max_aware = iter >= kRelaxIters;
    const int qmi = max_aware ? quad_max_counts(cnt_in, n, i) : 0;
    for (int a = 0; a < k; ++a) {
        const int32_t j = A[a];
        if (j == (int32_t)i || (j >> 2) == (int32_t)(i >> 2)) continue;
        bool first = true;
        for (int b = 0; b < a; ++b) first &= (A[b] != j);
        if (!first) continue;
        const int32_t* B = nbr + (int64_t)j * k;
        bool back = false;
        for (int b = 0; b < k; ++b) back |= (B[b] == (int32_t)i);
        if (!back) continue;                                           // one-sided pair: fixed owner
        const bool i_owns = (mine_old >> a) & 1ull;
        const bool in_tile = (j / kTileSpots) == (int32_t)(i / kTileSpots);     // the same kind for both endpoints
        const int sh = in_tile ? 0 : 16, one = 1 << sh;
        const int cii = (ci >> sh) & 0xffff, cj = (cnt_in[j] >> sh) & 0xffff;
        const int c_owner = i_owns ? cii : cj, c_other = i_owns ? cj : cii;
        bool flip;
        if (combined && max_aware) {
            // combined lists (an in-tile entry can sit in an out-of-tile slot): balance the totals, keep every
            // out-of-tile list within `combined` entries
            const int cnt_j = cnt_in[j];
            const int ti = (ci & 0xffff) + ((ci >> 16) & 0xffff), tj = (cnt_j & 0xffff) + ((cnt_j >> 16) & 0xffff);
            const int t_owner = i_owns ? ti : tj, t_other = i_owns ? tj : ti;
            const int mi = quad_max_total(cnt_in, n, i), mj = quad_max_total(cnt_in, n, j);
            const int m_owner = i_owns ? mi : mj, m_other = i_owns ? mj : mi;
            const int g_other = ((i_owns ? cnt_j : ci) >> 16) & 0xffff;
            uint32_t h = pair_hash((uint32_t)i, (uint32_t)j) ^ ((uint32_t)iter * 0x9E3779B9u);
            h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15;
            flip = ((t_owner > t_other + 1) || (t_owner == m_owner && t_owner > t_other && t_other + 1 <= m_other)) &&
                   (in_tile || g_other + 1 <= combined) && ((h & 3u) == 0u);
        } else if (!max_aware) {
            flip = (c_owner > c_other + 1) && ((pair_hash((uint32_t)i, (uint32_t)j) >> (iter + 1)) & 1u);
        } else {
            const int qmj = quad_max_counts(cnt_in, n, j);
            const int m_i = (qmi >> sh) & 0xffff, m_j = (qmj >> sh) & 0xffff;
            const int m_owner = i_owns ? m_i : m_j, m_other = i_owns ? m_j : m_i;
            uint32_t h = pair_hash((uint32_t)i, (uint32_t)j) ^ ((uint32_t)iter * 0x9E3779B9u);
            h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15;
            flip = ((c_owner > c_other + 1) || (c_owner == m_owner && c_owner > c_other && c_other + 1 <= m_other)) && ((h & 3u) == 0u);
        }
        if (flip) {
            if (i_owns) { bits &= ~(1ull << a); c -= one; }
            else { bits |= 1ull << a; c += one; }
        }
    }
    own_out[i] = bits;
    cnt_out[i] = c;
}

// walks the list entries spot i owns: f(j, coefficient, in_tile)
template <typename F>
__device__ __forceinline__ void plan_for_each_entry(const int32_t* __restrict__ nbr, int k, int64_t i, uint64_t bits, F f) {
    const int32_t* A = nbr + i * k;
    for (int a = 0; a < k; ++a) {
        const int32_t j = A[a];
        bool first = true;
        int mult = 0;
        for (int b = 0; b < k; ++b)
            if (A[b] == j) { if (b < a) first = false; ++mult; }
        if (!first) continue;
        if (j == (int32_t)i) { f(j, (float)mult, true); continue; }     // a_ii z_i^2: own row with coefficient a_ii
        if ((j >> 2) == (int32_t)(i >> 2)) continue;
        if (!((bits >> a) & 1ull)) continue;
        const int32_t* B = nbr + (int64_t)j * k;
        int back = 0;
        for (int b = 0; b < k; ++b) back += (B[b] == (int32_t)i);
#ifdef CHR_EXPERIMENT_DROP_GLOBAL   /* timing experiment only: wrong statistics */
        if ((j / kTileSpots) != (int32_t)(i / kTileSpots)) continue;
#endif
        f(j, (float)(mult + back), (j / kTileSpots) == (int32_t)(i / kTileSpots));
    }
}

// record size of every quad in 16-byte units.  R_glob = longest out-of-tile list of its spots; an in-tile entry
// may also sit in a free out-of-tile slot of its spot (addressed by its absolute row like the others), so
// R_loc + R_glob only has to cover the longest combined list.
__global__ void plan_size_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int64_t nquads, const uint64_t* __restrict__ own,
                                 int32_t* __restrict__ qsize, int32_t* __restrict__ qrounds) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquads) return;
    int rt = 0, rg = 0;
    for (int t = 0; t < 4; ++t) {
        const int64_t i = 4 * q + t;
        if (i >= n) break;
        int cl = 0, cg = 0;
        plan_for_each_entry(nbr, k, i, own[i], [&](int32_t, float, bool in_tile) { if (in_tile) ++cl; else ++cg; });
        rt = cl + cg > rt ? cl + cg : rt;
        rg = cg > rg ? cg : rg;
    }
    const int rl = rt - rg;                                             // >= 0: rt >= every cg
    qrounds[q] = rl | (rg << 8);
    qsize[q] = kPlanHdrUnits + kPlanRoundUnits * (rl + rg);
}

// single-CTA exclusive scan of m ints; out[m .. m + pad) = total
__global__ void plan_scan_kernel(const int32_t* __restrict__ in, int64_t m, int32_t* __restrict__ out, int pad) {
    __shared__ int64_t sh[1024];
    __shared__ int64_t total;
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
        total = run;
    }
    __syncthreads();
    int64_t run = sh[t];
    for (int64_t i = b; i < e; ++i) { out[i] = (int32_t)run; run += in[i]; }
    for (int i = t; i < pad; i += nt) out[m + i] = (int32_t)total;
}

// largest tile (16-byte units) and the plan header
__global__ void plan_header_kernel(const int32_t* __restrict__ off, int64_t n, int k, int64_t nquads, PlanHeader* __restrict__ hdr,
                                   int64_t off_bytes, int64_t rec_bytes) {
    __shared__ int smax[256];
    int m = 0;
    const int64_t ntiles = (nquads + kTileQuads - 1) / kTileQuads;
    for (int64_t c = threadIdx.x; c < ntiles; c += blockDim.x) {
        const int64_t q0 = c * kTileQuads, q1 = q0 + kTileQuads < nquads ? q0 + kTileQuads : nquads;
        const int u = off[q1] - off[q0];
        m = u > m ? u : m;
    }
    smax[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < (int)blockDim.x; ++t) m = smax[t] > m ? smax[t] : m;
        hdr->magic = kPlanMagic;
        hdr->k = k;
        hdr->n = n;
        hdr->nquads = nquads;
        hdr->chunk_units_max = m;
        hdr->total_units = off[nquads];
        hdr->off_bytes = off_bytes;
        hdr->rec_bytes = rec_bytes;
    }
}

// one thread per quad writes its record
__global__ void plan_fill_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int64_t nquads, const uint64_t* __restrict__ own,
                                 const int32_t* __restrict__ indeg, const int32_t* __restrict__ off, const int32_t* __restrict__ qrounds,
                                 int32_t* __restrict__ rec_base) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquads) return;
    int32_t* rec = rec_base + (int64_t)off[q] * 4;
    float* recf = reinterpret_cast<float*>(rec);
    const int RL = qrounds[q] & 0xff, RG = qrounds[q] >> 8;
    const int nvalid = (int)(n - 4 * q < 4 ? n - 4 * q : 4);
    for (int w = 1; w < kPlanHdrUnits * 4; ++w) recf[w] = 0.f;
    int irregular = 0;
    const int tile_row0 = (int)((4 * q) % kTileSpots);                 // row of the quad's first spot inside its tile
    int32_t* glob = rec + kPlanHdrUnits * 4;                           // out-of-tile rounds first (fixed position), then in-tile
    int32_t* loc = glob + RG * kPlanRoundUnits * 4;
    for (int r = 0; r < RL; ++r)
        for (int t = 0; t < 4; ++t) {
            loc[r * 8 + t] = (int32_t)0x80000000;                      // padding: negative offset (load skipped), coefficient 0
            loc[r * 8 + 4 + t] = 0;
        }
    for (int r = 0; r < RG; ++r)
        for (int t = 0; t < 4; ++t) {
            glob[r * 8 + t] = (int32_t)(4 * q + (t < nvalid ? t : 0));
            glob[r * 8 + 4 + t] = 0;
        }
    for (int t = 0; t < nvalid; ++t) {
        const int64_t i = 4 * q + t;
        int sl = 0, sg = 0;
        float csum = 0.f;
        plan_for_each_entry(nbr, k, i, own[i], [&](int32_t j, float c, bool in_tile) {
            if (in_tile) {
                if (sl < RL) { loc[sl * 8 + t] = (j % kTileSpots) * (kGeneTile * 4); reinterpret_cast<float*>(loc)[sl * 8 + 4 + t] = c; }
                ++sl;
            } else {
                if (sg < RG) { glob[sg * 8 + t] = j; reinterpret_cast<float*>(glob)[sg * 8 + 4 + t] = c; }
                ++sg;
            }
            csum += c;
        });
        if (sl > RL) {                                                  // in-tile entries beyond R_loc: free out-of-tile slots
            int seen = 0;
            plan_for_each_entry(nbr, k, i, own[i], [&](int32_t j, float c, bool in_tile) {
                if (!in_tile) return;
                if (seen++ >= RL && sg < RG) { glob[sg * 8 + t] = j; reinterpret_cast<float*>(glob)[sg * 8 + 4 + t] = c; ++sg; }
            });
        }
        recf[8 + t] = csum;
        recf[12 + t] = (float)(indeg[i] - k);                               // column sum of the adjacency minus its row sum
        if (indeg[i] != k) irregular = 1;
        // pairs inside the quad
        const int32_t* A = nbr + i * k;
        for (int a = 0; a < k; ++a) {
            const int32_t j = A[a];
            if (j == (int32_t)i || (j >> 2) != (int32_t)(i >> 2)) continue;
            int mult = 0, back = 0;
            for (int b = 0; b < k; ++b) mult += (A[b] == j);
            const int32_t* B = nbr + (int64_t)j * k;
            for (int b = 0; b < k; ++b) back += (B[b] == (int32_t)i);
            const int u = t, v = (int)(j & 3);
            const int lo = u < v ? u : v, hi = u < v ? v : u;
            const int idx = lo == 0 ? hi - 1 : (lo == 1 ? hi + 1 : 5);   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
            recf[1 + idx] = (float)(mult + back);                            // words 1..3 (unit 0), 4..6 (unit 1)
        }
    }
    // bit 30: some spot of the quad has indeg != k (the kernel then adds the (indeg - k) x terms)
    rec[0] = RL | (RG << 8) | (nvalid << 16) | (irregular << 30);
}

struct PlanScratch {
    size_t own[2], cnt[2], indeg, qsize, qrounds, total;
};
static PlanScratch plan_scratch(int64_t n) {
    PlanScratch p;
    const int64_t nquads = (n + 3) / 4;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += round_up<size_t>(b, 256); return r; };
    for (int b = 0; b < 2; ++b) { p.own[b] = take((size_t)n * 8); p.cnt[b] = take((size_t)n * 4); }
    p.indeg = take((size_t)n * 4);
    p.qsize = take((size_t)nquads * 4);
    p.qrounds = take((size_t)nquads * 4);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

size_t chr_moran_plan_bytes(int64_t n_spots, int k) {
    if (n_spots <= 0 || k <= 0) return chr::kPlanHeaderBytes;
    const int64_t nquads = (n_spots + 3) / 4;
    return chr::kPlanHeaderBytes + chr::round_up<size_t>((size_t)(nquads + 1 + chr::kPlanOffPad) * 4, 256) +
           chr::plan_rec_capacity_bytes(n_spots, k) + 256;
}

size_t chr_moran_plan_workspace_bytes(int64_t n_spots, int k) {
    (void)k;
    return chr::plan_scratch(n_spots > 0 ? n_spots : 1).total + 512;
}

int chr_moran_plan_build(const int32_t* nbr, int64_t n_spots, int k, void* plan, size_t plan_bytes, int32_t* info_host,
                         void* workspace, size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(nbr && plan && workspace, "chr_moran_plan_build: null pointer argument");
    CHR_REQUIRE(n_spots > 1 && n_spots < (1LL << 31), "chr_moran_plan_build: need 1 < n_spots < 2^31 (got %lld)", (long long)n_spots);
    CHR_REQUIRE(k > 0 && k <= 64 && k < n_spots, "chr_moran_plan_build: need 0 < k <= 64 and k < n_spots (k=%d)", k);
    CHR_REQUIRE(((uintptr_t)plan & 255) == 0, "chr_moran_plan_build: plan must be 256-byte aligned");
    CHR_REQUIRE(plan_bytes >= chr_moran_plan_bytes(n_spots, k), "chr_moran_plan_build: plan buffer too small (%zu < %zu)", plan_bytes,
                chr_moran_plan_bytes(n_spots, k));
    PlanScratch sc = plan_scratch(n_spots);
    CHR_REQUIRE(workspace_bytes >= sc.total + 256, "chr_moran_plan_build: workspace too small (%zu < %zu)", workspace_bytes, sc.total + 256);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint64_t* own[2] = {(uint64_t*)(ws + sc.own[0]), (uint64_t*)(ws + sc.own[1])};
    int32_t* cnt[2] = {(int32_t*)(ws + sc.cnt[0]), (int32_t*)(ws + sc.cnt[1])};
    int32_t* indeg = (int32_t*)(ws + sc.indeg);
    int32_t* qsize = (int32_t*)(ws + sc.qsize);
    int32_t* qrounds = (int32_t*)(ws + sc.qrounds);
    const int64_t nquads = (n_spots + 3) / 4;
    const int64_t off_bytes = kPlanHeaderBytes;
    const int64_t rec_bytes = off_bytes + (int64_t)round_up<size_t>((size_t)(nquads + 1 + kPlanOffPad) * 4, 256);
    PlanHeader* hdr = (PlanHeader*)plan;
    int32_t* off = (int32_t*)((char*)plan + off_bytes);
    int32_t* rec = (int32_t*)((char*)plan + rec_bytes);

    const unsigned gs = (unsigned)ceil_div<int64_t>(n_spots, 128), gq = (unsigned)ceil_div<int64_t>(nquads, 128);
    timer_begin(CHR_FAMILY_KNN, timing_env(), st);
    CHR_CUDA(cudaMemsetAsync(indeg, 0, (size_t)n_spots * 4, st));
    plan_init_kernel<<<gs, 128, 0, st>>>(nbr, n_spots, k, own[0], cnt[0], indeg);
    int cur = 0;
    const char* env_sweeps = getenv("CHR_PLAN_SWEEPS");   /* experiment hook */
    const int n_sweeps = env_sweeps ? atoi(env_sweeps) : kRelaxIters + kRelaxMaxIters;
    const char* env_mode = getenv("CHR_PLAN_COMBINED");     /* experiment hook */
    const int combined = env_mode ? atoi(env_mode) : 0;
    for (int it = 0; it < n_sweeps; ++it, cur ^= 1)
        plan_relax_kernel<<<gs, 128, 0, st>>>(nbr, n_spots, k, it, combined, own[cur], cnt[cur], own[cur ^ 1], cnt[cur ^ 1]);
    plan_size_kernel<<<gq, 128, 0, st>>>(nbr, n_spots, k, nquads, own[cur], qsize, qrounds);
    plan_scan_kernel<<<1, 1024, 0, st>>>(qsize, nquads, off, 1 + kPlanOffPad);
    plan_header_kernel<<<1, 256, 0, st>>>(off, n_spots, k, nquads, hdr, off_bytes, rec_bytes);
    plan_fill_kernel<<<gq, 128, 0, st>>>(nbr, n_spots, k, nquads, own[cur], indeg, off, qrounds, rec);
    timer_end(CHR_FAMILY_KNN, timing_env(), st);
    CHR_LAUNCH_CHECK();
    if (info_host) {
        PlanHeader h;
        CHR_CUDA(cudaMemcpyAsync(&h, hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
        CHR_CUDA(cudaStreamSynchronize(st));
        info_host[0] = h.chunk_units_max;
        info_host[1] = h.total_units;
        info_host[2] = (int32_t)nquads;
        info_host[3] = 0;
    }
    return 0;
}

}  // extern "C"
