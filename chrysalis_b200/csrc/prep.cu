// K2 - preprocessing front-end on CSR counts (spots x genes).
//
// Replaces  sc.pp.filter_genes(min_cells) / sc.pp.normalize_total / sc.pp.log1p / ad.to_df()
// of /root/reference/chrysalis/core.py:53-59:
//   gene_nnz[g]   = #{i : X[i,g] > 0}                      (filter_genes keeps gene_nnz >= min_cells)
//   totals[i]     = sum over KEPT genes of X[i,g]          (normalize_total runs on the filtered copy)
//   dense[r(i)][c(g)] = log1p(X[i,g] * scale[i])           (scale = median(totals>0)/totals, host side)
// The dense float32 matrix is the reference's `gene_matrix = ad.to_df()`; rows can be written
// through a permutation so the matrix comes out in Hilbert (spatially coherent) spot order.
#include "common.cuh"

namespace chr {

// nnz is read from indptr[n] on the device, so the launch needs no host sync
__global__ void csr_gene_nnz_rows_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                         const float* __restrict__ data, int64_t n, int32_t* __restrict__ gene_nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t b = indptr[i], e = indptr[i + 1];
        for (int64_t t = b + lane; t < e; t += 32)
            if (data[t] > 0.f) atomicAdd(gene_nnz + indices[t], 1);
    }
}

// privatised variant: one histogram per CTA in shared memory (n_genes * 4 bytes), flat coalesced sweep over the
// stored entries (nnz = indptr[n], read on the device), merged with one global atomic per gene and CTA
__global__ void __launch_bounds__(1024)
csr_gene_nnz_smem_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ data,
                         int64_t n, int n_genes, int32_t* __restrict__ gene_nnz) {
    extern __shared__ int32_t hist[];
    for (int g = threadIdx.x; g < n_genes; g += blockDim.x) hist[g] = 0;
    __syncthreads();
    const int64_t nnz = indptr[n];
    const int64_t nvec = nnz >> 2;                                  // both arrays are 16-byte aligned (checked by the caller)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int4* __restrict__ idx4 = reinterpret_cast<const int4*>(indices);
    const float4* __restrict__ dat4 = reinterpret_cast<const float4*>(data);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int4 c = __ldcs(idx4 + v);
        const float4 x = __ldcs(dat4 + v);
        if (x.x > 0.f) atomicAdd(hist + c.x, 1);
        if (x.y > 0.f) atomicAdd(hist + c.y, 1);
        if (x.z > 0.f) atomicAdd(hist + c.z, 1);
        if (x.w > 0.f) atomicAdd(hist + c.w, 1);
    }
    if (blockIdx.x == 0 && threadIdx.x < (nnz & 3)) {
        const int64_t e = (nvec << 2) + threadIdx.x;
        if (data[e] > 0.f) atomicAdd(hist + indices[e], 1);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
        const int32_t h = hist[g];
        if (h) atomicAdd(gene_nnz + g, h);
    }
}

__global__ void csr_spot_totals_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_col,
                                       double* __restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t b = indptr[i], e = indptr[i + 1];
        double s = 0.0;
        for (int64_t t = b + lane; t < e; t += 32)
            if (gene_col[indices[t]] >= 0) s += (double)data[t];
        s = warp_sum(s);
        if (lane == 0) totals[i] = s;
    }
}

__global__ void csr_densify_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_col,
                                   const double* __restrict__ scale, const int64_t* __restrict__ row_of_spot,
                                   float* __restrict__ dense, int64_t ld, int apply_log1p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarp) {
        const int64_t b = indptr[i], e = indptr[i + 1];
        const double sc = scale ? scale[i] : 1.0;
        float* __restrict__ row = dense + (row_of_spot ? row_of_spot[i] : i) * ld;
        for (int64_t t = b + lane; t < e; t += 32) {
            const int c = gene_col[indices[t]];
            if (c < 0) continue;
            float v = scale ? (float)((double)data[t] * sc) : data[t];
            if (apply_log1p) v = log1pf(v);
            row[c] = v;
        }
    }
}

// compact counts -> the int32 / float32 arrays the other kernels read: u16 column indices, u8 values (255 = escape)
__global__ void csr_decode_compact_kernel(const uint16_t* __restrict__ cols, const uint8_t* __restrict__ vals, int64_t nnz,
                                          int32_t* __restrict__ out_idx, float* __restrict__ out_val) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; t < nnz; t += stride) {
        if (t + 8 <= nnz) {                                           // 16 bytes of columns + 8 bytes of values per thread
            const uint4 c = *reinterpret_cast<const uint4*>(cols + t);
            const uint2 v = *reinterpret_cast<const uint2*>(vals + t);
            int4 i0 = make_int4(c.x & 0xffff, c.x >> 16, c.y & 0xffff, c.y >> 16), i1 = make_int4(c.z & 0xffff, c.z >> 16, c.w & 0xffff, c.w >> 16);
            float4 f0 = make_float4((float)(v.x & 0xff), (float)((v.x >> 8) & 0xff), (float)((v.x >> 16) & 0xff), (float)(v.x >> 24));
            float4 f1 = make_float4((float)(v.y & 0xff), (float)((v.y >> 8) & 0xff), (float)((v.y >> 16) & 0xff), (float)(v.y >> 24));
            *reinterpret_cast<int4*>(out_idx + t) = i0;
            *reinterpret_cast<int4*>(out_idx + t + 4) = i1;
            *reinterpret_cast<float4*>(out_val + t) = f0;
            *reinterpret_cast<float4*>(out_val + t + 4) = f1;
        } else {
            for (int64_t u = t; u < nnz; ++u) { out_idx[u] = cols[u]; out_val[u] = (float)vals[u]; }
        }
    }
}
__global__ void csr_decode_overflow_kernel(const int64_t* __restrict__ pos, const float* __restrict__ val, int64_t m, float* __restrict__ out_val) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) out_val[pos[t]] = val[t];
}

// delta-coded columns: a stored entry's column = the previous entry's column of the same row + delta (the first entry of a row
// counts from 0); delta 255 = "look up the true delta" (gap_pos sorted positions, gap_val).  One warp per row: 32 deltas per
// step, warp-inclusive scan, running base.
__global__ void __launch_bounds__(256)
csr_decode_delta_kernel(const int64_t* __restrict__ indptr, const uint8_t* __restrict__ deltas, const uint8_t* __restrict__ vals, int64_t n_rows,
                        const int64_t* __restrict__ gap_pos, const int32_t* __restrict__ gap_val, int64_t n_gap,
                        int32_t* __restrict__ out_idx, float* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += n_warps) {
        const int64_t p0 = indptr[r], p1 = indptr[r + 1];
        int base = 0;
        for (int64_t q = p0; q < p1; q += 32) {
            const int64_t p = q + lane;
            int dlt = 0;
            float v = 0.f;
            if (p < p1) {
                dlt = deltas[p];
                v = (float)vals[p];
                if (dlt == 255) {                                     // rare: binary search of the position
                    int64_t lo = 0, hi = n_gap - 1;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (gap_pos[mid] < p) lo = mid + 1; else hi = mid;
                    }
                    dlt = gap_val[lo];
                }
            }
            int incl = dlt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (p < p1) {
                out_idx[p] = base + incl;
                out_val[p] = v;
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

}  // namespace chr

extern "C" {

int chr_csr_decode_compact_delta(const int64_t* indptr, const uint8_t* deltas, const uint8_t* vals, int64_t n_rows, int64_t nnz,
                                 const int64_t* gap_pos, const int32_t* gap_val, int64_t n_gap, const int64_t* over_pos,
                                 const float* over_val, int64_t n_over, int32_t* out_indices, float* out_data, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(n_rows > 0 && nnz >= 0 && n_gap >= 0 && n_over >= 0 && indptr && (nnz == 0 || (deltas && vals && out_indices && out_data)) &&
                    (n_gap == 0 || (gap_pos && gap_val)) && (n_over == 0 || (over_pos && over_val)),
                "chr_csr_decode_compact_delta: bad arguments");
    if (nnz > 0) csr_decode_delta_kernel<<<kNumSMs * 16, 256, 0, st>>>(indptr, deltas, vals, n_rows, gap_pos, gap_val, n_gap, out_indices, out_data);
    if (n_over > 0) csr_decode_overflow_kernel<<<(unsigned)ceil_div<int64_t>(n_over, 256), 256, 0, st>>>(over_pos, over_val, n_over, out_data);
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_csr_decode_compact(const uint16_t* cols, const uint8_t* vals, int64_t nnz, const int64_t* over_pos, const float* over_val,
                           int64_t n_over, int32_t* out_indices, float* out_data, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(nnz >= 0 && n_over >= 0 && (nnz == 0 || (cols && vals && out_indices && out_data)) && (n_over == 0 || (over_pos && over_val)),
                "chr_csr_decode_compact: bad arguments");
    CHR_REQUIRE((((uintptr_t)cols | (uintptr_t)out_indices | (uintptr_t)out_data) & 15) == 0 && ((uintptr_t)vals & 7) == 0,
                "chr_csr_decode_compact: arrays must be 16-byte aligned (values 8-byte)");
    if (nnz > 0) csr_decode_compact_kernel<<<kNumSMs * 8, 256, 0, st>>>(cols, vals, nnz, out_indices, out_data);
    if (n_over > 0) csr_decode_overflow_kernel<<<(unsigned)ceil_div<int64_t>(n_over, 256), 256, 0, st>>>(over_pos, over_val, n_over, out_data);
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_csr_gene_nnz(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n, int64_t n_genes,
                     int32_t* gene_nnz, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(indptr && indices && data && gene_nnz && n > 0 && n_genes > 0, "chr_csr_gene_nnz: bad arguments");
    CHR_CUDA(cudaMemsetAsync(gene_nnz, 0, (size_t)n_genes * sizeof(int32_t), st));
    const size_t hist_bytes = (size_t)n_genes * sizeof(int32_t);
    const bool aligned = (((uintptr_t)indices | (uintptr_t)data) & 15) == 0;
    if (aligned && hist_bytes <= 200 * 1024) {
        if (hist_bytes > 48 * 1024)
            CHR_CUDA(cudaFuncSetAttribute(csr_gene_nnz_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
        csr_gene_nnz_smem_kernel<<<kNumSMs, 1024, hist_bytes, st>>>(indptr, indices, data, n, (int)n_genes, gene_nnz);
    } else {
        csr_gene_nnz_rows_kernel<<<kNumSMs * 8, 256, 0, st>>>(indptr, indices, data, n, gene_nnz);
    }
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_csr_spot_totals(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                        const int32_t* gene_col, double* totals, chr_stream_t stream) {
    using namespace chr;
    CHR_REQUIRE(indptr && indices && data && gene_col && totals && n > 0, "chr_csr_spot_totals: bad arguments");
    csr_spot_totals_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, n, gene_col, totals);
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_csr_densify_log1p(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                          const int32_t* gene_col, const double* scale, const int64_t* row_of_spot, float* dense,
                          int64_t ld, int apply_log1p, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(indptr && indices && data && gene_col && dense && n > 0 && ld > 0, "chr_csr_densify_log1p: bad arguments");
    if (!(apply_log1p & 2)) CHR_CUDA(cudaMemsetAsync(dense, 0, (size_t)n * ld * sizeof(float), st));   // bit 1: caller zeroed it
    timer_begin(CHR_FAMILY_PREP, timing_env(), st);
    csr_densify_kernel<<<kNumSMs * 8, 256, 0, st>>>(indptr, indices, data, n, gene_col, scale, row_of_spot, dense, ld,
                                                    apply_log1p & 1);
    timer_end(CHR_FAMILY_PREP, timing_env(), st);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
