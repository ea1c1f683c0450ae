// K1 - spatial weights: exact k-nearest-neighbour lists on the GPU (uniform-grid binning).
//
// Replaces  weights.KNN.from_array(points, k); w.transform = 'R'
// (/root/reference/chrysalis/core.py:64-65; libpysal KNN = scipy cKDTree.query(points, k+1)
// with the query point dropped).  Row-standardisation is implicit (every weight 1/k), so the
// CSR of W is the ELL list nbr[n][k] itself: indptr = k*i, data = 1/k.
//
// Distances are float64 like the reference's.  Ties are broken by the lower spot index
// (cKDTree's tie order depends on tree traversal and is not reproducible - SURVEY.md H2).
#include "common.cuh"

namespace chr {

constexpr int kKnnMaxK = 64;

struct KnnGrid {
    double x0, y0, cell, inv_cell;
    int gx, gy;
};

// single-CTA bounding box (n is at most a few million points; runs once per sample)
__global__ void bbox_kernel(const double* __restrict__ xy, int64_t n, double* __restrict__ bbox) {
    double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
    __shared__ double sh[4][32];
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { sh[0][warp] = mnx; sh[1][warp] = mny; sh[2][warp] = mxx; sh[3][warp] = mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) {
            mnx = fmin(mnx, sh[0][w]); mny = fmin(mny, sh[1][w]);
            mxx = fmax(mxx, sh[2][w]); mxy = fmax(mxy, sh[3][w]);
        }
        bbox[0] = mnx; bbox[1] = mny; bbox[2] = mxx; bbox[3] = mxy;
    }
}

__device__ __forceinline__ KnnGrid make_grid(const double* __restrict__ bbox, int64_t n) {
    KnnGrid g;
    g.x0 = bbox[0]; g.y0 = bbox[1];
    double w = bbox[2] - bbox[0], h = bbox[3] - bbox[1];
    if (!(w > 0)) w = 1.0;
    if (!(h > 0)) h = 1.0;
    // ~2 spots per cell on average, grid capped at 2048 x 2048 cells
    double cell = sqrt(w * h * 2.0 / (double)n);
    cell = fmax(cell, fmax(w, h) / 2047.0);
    g.cell = cell;
    g.inv_cell = 1.0 / cell;
    g.gx = (int)(w * g.inv_cell) + 1;
    g.gy = (int)(h * g.inv_cell) + 1;
    return g;
}

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_cell, int gmax) {
    int c = (int)((v - v0) * inv_cell);
    return c < 0 ? 0 : (c >= gmax ? gmax - 1 : c);
}

__global__ void knn_count_kernel(const double* __restrict__ xy, int64_t n, const double* __restrict__ bbox,
                                 int32_t* __restrict__ cell_of, int32_t* __restrict__ cell_cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    KnnGrid g = make_grid(bbox, n);
    int cx = cell_coord(xy[2 * i], g.x0, g.inv_cell, g.gx), cy = cell_coord(xy[2 * i + 1], g.y0, g.inv_cell, g.gy);
    int c = cy * g.gx + cx;
    cell_of[i] = c;
    atomicAdd(cell_cnt + c, 1);
}

// single-CTA exclusive scan of m ints (m <= a few million); out[m] = total
__global__ void exclusive_scan_kernel(const int32_t* __restrict__ in, int64_t m, int32_t* __restrict__ out) {
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

__global__ void knn_fill_kernel(int64_t n, const int32_t* __restrict__ cell_of, const int32_t* __restrict__ cell_start,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ sorted_ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    int slot = atomicAdd(cursor + c, 1);
    sorted_ids[cell_start[c] + slot] = (int32_t)i;
}

// one thread per spot: expanding square rings of cells until the k-th best is provably final
__global__ void __launch_bounds__(128)
knn_query_kernel(const double* __restrict__ xy, int64_t n, int k, const double* __restrict__ bbox,
                 const int32_t* __restrict__ cell_start, const int32_t* __restrict__ sorted_ids,
                 int32_t* __restrict__ nbr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    KnnGrid g = make_grid(bbox, n);
    const double px = xy[2 * i], py = xy[2 * i + 1];
    const int cx = cell_coord(px, g.x0, g.inv_cell, g.gx), cy = cell_coord(py, g.y0, g.inv_cell, g.gy);

    double bd[kKnnMaxK];
    int32_t bi[kKnnMaxK];
    int cnt = 0;

    auto visit_cell = [&](int ax, int ay) {
        const int c = ay * g.gx + ax;
        const int b = cell_start[c], e = cell_start[c + 1];
        for (int s = b; s < e; ++s) {
            const int32_t j = sorted_ids[s];
            if (j == (int32_t)i) continue;
            const double dx = xy[2 * (int64_t)j] - px, dy = xy[2 * (int64_t)j + 1] - py;
            const double d = dx * dx + dy * dy;
            if (cnt == k && !(d < bd[k - 1] || (d == bd[k - 1] && j < bi[k - 1]))) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && (bd[pos - 1] > d || (bd[pos - 1] == d && bi[pos - 1] > j))) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = j;
            if (cnt < k) ++cnt;
        }
    };

    const int rmax = g.gx > g.gy ? g.gx : g.gy;
    for (int r = 0; r <= rmax; ++r) {
        const int x_lo = cx - r, x_hi = cx + r, y_lo = cy - r, y_hi = cy + r;
        if (r == 0) {
            visit_cell(cx, cy);
        } else {
            for (int ax = max(x_lo, 0); ax <= min(x_hi, g.gx - 1); ++ax) {
                if (y_lo >= 0) visit_cell(ax, y_lo);
                if (y_hi < g.gy) visit_cell(ax, y_hi);
            }
            for (int ay = max(y_lo + 1, 0); ay <= min(y_hi - 1, g.gy - 1); ++ay) {
                if (x_lo >= 0) visit_cell(x_lo, ay);
                if (x_hi < g.gx) visit_cell(x_hi, ay);
            }
        }
        // sides of the visited square that lie on/outside the grid have no points beyond them
        const double INF = 1e300;
        const double sx_lo = x_lo <= 0 ? INF : px - (g.x0 + x_lo * g.cell);
        const double sx_hi = x_hi >= g.gx - 1 ? INF : (g.x0 + (x_hi + 1) * g.cell) - px;
        const double sy_lo = y_lo <= 0 ? INF : py - (g.y0 + y_lo * g.cell);
        const double sy_hi = y_hi >= g.gy - 1 ? INF : (g.y0 + (y_hi + 1) * g.cell) - py;
        double safe = fmin(fmin(sx_lo, sx_hi), fmin(sy_lo, sy_hi));
        if (safe >= INF) break;  // whole grid visited
        // every unvisited point is at least `safe` away (minus cell-assignment rounding slack)
        safe -= 1e-9 * g.cell;
        if (cnt == k && safe > 0 && bd[k - 1] < safe * safe) break;
    }
    for (int j = 0; j < k; ++j) nbr[i * k + j] = j < cnt ? bi[j] : 0;
}

// ---- Hilbert keys -----------------------------------------------------------------------
__global__ void hilbert_kernel(const double* __restrict__ xy, int64_t n, const double* __restrict__ bbox,
                               int64_t* __restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w = bbox[2] - bbox[0], h = bbox[3] - bbox[1];
    double span = fmax(fmax(w, h), 1e-300);
    const double sc = 65535.0 / span;
    uint32_t x = (uint32_t)fmin(fmax((xy[2 * i] - bbox[0]) * sc, 0.0), 65535.0);
    uint32_t y = (uint32_t)fmin(fmax((xy[2 * i + 1] - bbox[1]) * sc, 0.0), 65535.0);
    uint64_t d = 0;
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            uint32_t t = x; x = y; y = t;
        }
    }
    keys[i] = (int64_t)d;
}

struct KnnPlan {
    int64_t max_cells;
    size_t off_cell_of, off_cnt, off_start, off_cursor, off_sorted, total;
};

static KnnPlan knn_plan(int64_t n) {
    KnnPlan p;
    // gx*gy = wh/cell^2 + (w+h)/cell + 1 <= n/2 + 2*2047 + 1 (cell >= max(w,h)/2047, cell^2 >= 2wh/n)
    p.max_cells = n / 2 + 4100;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += round_up<size_t>(bytes, 256); return r; };
    p.off_cell_of = take((size_t)n * 4);
    p.off_cnt = take((size_t)(p.max_cells + 1) * 4);
    p.off_start = take((size_t)(p.max_cells + 1) * 4);
    p.off_cursor = take((size_t)(p.max_cells + 1) * 4);
    p.off_sorted = take((size_t)n * 4);
    p.total = o;
    return p;
}

}  // namespace chr

extern "C" {

int chr_spatial_bbox(const double* xy, int64_t n, double* bbox_out, chr_stream_t stream) {
    CHR_REQUIRE(xy && bbox_out && n > 0, "chr_spatial_bbox: bad arguments");
    chr::bbox_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(xy, n, bbox_out);
    CHR_LAUNCH_CHECK();
    return 0;
}

size_t chr_knn_workspace_bytes(int64_t n, int k) {
    (void)k;
    return chr::knn_plan(n > 0 ? n : 1).total + 512;
}

int chr_knn_build(const double* xy, int64_t n, int k, const double* bbox, int32_t* nbr_out, void* workspace,
                  size_t workspace_bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(xy && bbox && nbr_out && workspace, "chr_knn_build: null pointer argument");
    CHR_REQUIRE(n > 1 && n < (1LL << 31), "chr_knn_build: n out of range (%lld)", (long long)n);
    CHR_REQUIRE(k > 0 && k < n && k <= kKnnMaxK, "chr_knn_build: need 0 < k < n and k <= %d (k=%d)", kKnnMaxK, k);
    KnnPlan p = knn_plan(n);
    CHR_REQUIRE(workspace_bytes >= p.total, "chr_knn_build: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* cell_of = (int32_t*)(ws + p.off_cell_of);
    int32_t* cell_cnt = (int32_t*)(ws + p.off_cnt);
    int32_t* cell_start = (int32_t*)(ws + p.off_start);
    int32_t* cursor = (int32_t*)(ws + p.off_cursor);
    int32_t* sorted_ids = (int32_t*)(ws + p.off_sorted);

    CHR_CUDA(cudaMemsetAsync(cell_cnt, 0, (size_t)(p.max_cells + 1) * 4, st));
    CHR_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(p.max_cells + 1) * 4, st));
    const unsigned nb = (unsigned)ceil_div<int64_t>(n, 256);
    knn_count_kernel<<<nb, 256, 0, st>>>(xy, n, bbox, cell_of, cell_cnt);
    CHR_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(cell_cnt, p.max_cells, cell_start);
    CHR_LAUNCH_CHECK();
    knn_fill_kernel<<<nb, 256, 0, st>>>(n, cell_of, cell_start, cursor, sorted_ids);
    CHR_LAUNCH_CHECK();
    timer_begin(CHR_FAMILY_KNN, timing_env(), st);
    knn_query_kernel<<<(unsigned)ceil_div<int64_t>(n, 128), 128, 0, st>>>(xy, n, k, bbox, cell_start, sorted_ids, nbr_out);
    timer_end(CHR_FAMILY_KNN, timing_env(), st);
    CHR_LAUNCH_CHECK();
    return 0;
}

int chr_hilbert_keys(const double* xy, int64_t n, const double* bbox, int64_t* keys_out, chr_stream_t stream) {
    CHR_REQUIRE(xy && bbox && keys_out && n > 0, "chr_hilbert_keys: bad arguments");
    chr::hilbert_kernel<<<(unsigned)chr::ceil_div<int64_t>(n, 256), 256, 0, (cudaStream_t)stream>>>(xy, n, bbox, keys_out);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
