// Shared definitions of the Moran's I kernels (moran.cu) and the W -> plan builder (moran_plan.cu).
#pragma once
#include "common.cuh"

namespace chr {

constexpr int kGeneTile = 128;    // genes per CTA = 32 lanes x float4 (one 512-byte row slab)
constexpr int kNStat = 5;         // S1, S2, SW, SC, S5 (Geary)
constexpr int kTileSpots = 112;   // spots per staged tile = 28 quads = 2 per compute warp (tile t = spots [112 t, 112 t + 112))
constexpr int kTileQuads = kTileSpots / 4;

// ---- quad plan of W (see moran_plan.cu) -------------------------------------------------------
// record of one quad, 16-byte units:
//   unit 0: meta (R_loc | R_glob << 8 | nvalid << 16), c01, c02, c03        coefficients of the pairs inside the quad
//   unit 1: c12, c13, c23, 0
//   unit 2: csum[4]   sum of the list coefficients of each spot (mean-shift term)
//   unit 3: indeg[4]  column sums of the adjacency (sum_j indeg_j x_j = 1'A x)
//   then R_loc rounds  { int off[4] (byte offset of the row inside the staged tile), float coef[4] }
//   then R_glob rounds { int row[4] (global row index),                              float coef[4] }
constexpr int kPlanHdrUnits = 4;
constexpr int kPlanRoundUnits = 2;
constexpr int kPlanOffPad = 64;          // ints: the offset table is readable kPlanOffPad entries past nquads
constexpr int kPlanHeaderBytes = 256;
constexpr uint32_t kPlanMagic = 0x51504c4eu;

struct PlanHeader {                      // first kPlanHeaderBytes of the plan buffer (device)
    uint32_t magic;
    int32_t k;
    int64_t n, nquads;
    int32_t chunk_units_max;             // largest number of record units of one tile
    int32_t total_units;
    int64_t off_bytes, rec_bytes;        // byte offsets of the offset table / records inside the plan buffer
};

inline size_t plan_rec_capacity_bytes(int64_t n, int k) {
    // every pair is owned once and a spot owns at most k pairs (+ a self entry): sum over quads of
    // (R_loc + R_glob) <= 2 * sum of list lengths
    const int64_t nquads = (n + 3) / 4;
    return (size_t)nquads * kPlanHdrUnits * 16 + (size_t)n * (size_t)(k + 1) * 2 * kPlanRoundUnits * 16;
}

// fixed-order fp64 sum of the per-CTA partials + the statistics (moran.cu); used by the lattice kernel as well
int launch_moran_finalize(const double* part, int nsplit, int64_t gpad, int64_t n_genes, int64_t n_spots, int k, const float* pilot,
                          bool geary, double* out_I, double* out_C, double* out_stats, cudaStream_t st);

// ---- packed f32x2 arithmetic --------------------------------------------------------------------
typedef unsigned long long f2_t;  // two packed fp32 values in a 64-bit register pair

__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t c; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) { f2_t c; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_pack(float x, float y) { f2_t c; asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(x), "f"(y)); return c; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f2_t f2_dup(float c) { return f2_pack(c, c); }

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
__device__ __forceinline__ void r4_fmac(Row4& s, float c, const Row4& v) {   // s += c * v
    const f2_t cc = f2_dup(c);
    s.a = f2_fma(cc, v.a, s.a);
    s.b = f2_fma(cc, v.b, s.b);
}
__device__ __forceinline__ void r4_sqdiff(Row4& q, const Row4& x, const Row4& v) {
    Row4 d = x;
    r4_sub(d, v);
    r4_fma(q, d, d);
}

}  // namespace chr
