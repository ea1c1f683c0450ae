// Host side of the boundary: getting an AnnData's CSR matrix from PAGEABLE host memory onto the device.
//
// The reference hands `adata.X` (scipy CSR in ordinary numpy arrays) to scanpy / pandas on the CPU
// (/root/reference/chrysalis/core.py:53-59 for detect_svgs, :126 for pca); here the same arrays have to cross PCIe.
// A plain cudaMemcpy from pageable memory is staged by the driver on one thread (~11 GB/s measured); these entry
// points fill a ring of pinned bounce buffers with a pool of host threads and overlap the fill of chunk c + 1 with
// the DMA of chunk c:
//   chr_host_upload            any buffer, as is
//   chr_host_download          the way back for the 10-100 MB results (PCA scores, AA proportions): DMA into the ring,
//                              pool threads copy (and first-touch) the caller's fresh pageable array
//   chr_host_csr_select_*      only the entries of selected columns (pca densifies the SVG columns, core.py:126):
//                              count pass + pack pass, both multi-threaded; ~1/9 of the bytes cross PCIe at cfg4
// The ring (4 x 32 MB, cudaHostAlloc'd once per process) and the thread pool are the only library-level state.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace chr {

class HostPool {
public:
    explicit HostPool(int n) : n_(n) {
        for (int t = 0; t < n_; ++t) th_.emplace_back([this, t] { loop(t); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    // runs f(t) for t = 0 .. size-1 on the pool threads and waits for all of them
    void run(const std::function<void(int)>& f) {
        std::unique_lock<std::mutex> l(m_);
        job_ = &f;
        pending_ = n_;
        ++gen_;
        cv_.notify_all();
        done_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int t) {
        int seen = 0;
        for (;;) {
            const std::function<void(int)>* f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = job_;
            }
            (*f)(t);
            {
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    int gen_ = 0, pending_ = 0;
    bool stop_ = false;
};

constexpr int kRingSlots = 4;
constexpr size_t kRingSlotBytes = (size_t)32 << 20;

struct HostIo {
    std::mutex mu;                       // one upload at a time per process (the ring is shared)
    HostPool* pool = nullptr;
    char* slot[kRingSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kRingSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[kRingSlots] = {false, false, false, false};
};
static HostIo g_io;

static int host_threads() {
    const char* e = getenv("CHR_HOST_THREADS");
    int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 32) n = 32;
    return n;
}

// lazily creates pool, ring and events; returns 0 on success (call with g_io.mu held)
static int host_io_ready() {
    if (!g_io.pool) g_io.pool = new HostPool(host_threads());
    for (int s = 0; s < kRingSlots; ++s) {
        if (!g_io.slot[s]) CHR_CUDA(cudaHostAlloc((void**)&g_io.slot[s], kRingSlotBytes, cudaHostAllocPortable));
        if (!g_io.ev[s]) CHR_CUDA(cudaEventCreateWithFlags(&g_io.ev[s], cudaEventDisableTiming));
    }
    return 0;
}

static int ring_acquire(int s) {
    if (g_io.busy[s]) {
        CHR_CUDA(cudaEventSynchronize(g_io.ev[s]));
        g_io.busy[s] = false;
    }
    return 0;
}
static int ring_submit(int s, void* dst, size_t bytes, cudaStream_t st) {
    if (bytes) CHR_CUDA(cudaMemcpyAsync(dst, g_io.slot[s], bytes, cudaMemcpyHostToDevice, st));
    CHR_CUDA(cudaEventRecord(g_io.ev[s], st));
    g_io.busy[s] = true;
    return 0;
}

}  // namespace chr

extern "C" {

int chr_host_threads(void) { return chr::host_threads(); }

// dst (device) <- src (pageable or pinned host), asynchronous on `stream` from the device's point of view; the host
// returns when the last chunk has been handed to the copy engine (the bounce buffers are recycled through events).
int chr_host_upload(void* dst_device, const void* src_host, size_t bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(bytes == 0 || (dst_device && src_host), "chr_host_upload: null pointer argument");
    std::lock_guard<std::mutex> lock(g_io.mu);
    if (int rc = host_io_ready()) return rc;
    const int nt = g_io.pool->size();
    int s = 0;
    for (size_t off = 0; off < bytes; off += kRingSlotBytes, s = (s + 1) % kRingSlots) {
        const size_t len = std::min(kRingSlotBytes, bytes - off);
        if (int rc = ring_acquire(s)) return rc;
        const char* src = (const char*)src_host + off;
        char* buf = g_io.slot[s];
        g_io.pool->run([&](int t) {
            const size_t per = ((len + nt - 1) / nt + 63) & ~(size_t)63;
            const size_t b = std::min(len, per * t), e = std::min(len, b + per);
            if (e > b) memcpy(buf + b, src + b, e - b);
        });
        if (int rc = ring_submit(s, (char*)dst_device + off, len, st)) return rc;
    }
    return 0;
}

// dst (pageable or pinned host) <- src (device).  Synchronous: returns when dst holds the data.  The copy of chunk c out of its
// bounce buffer (all pool threads; a fresh numpy array is first-touched here, in parallel) overlaps the DMA of chunk c + 1.
int chr_host_download(void* dst_host, const void* src_device, size_t bytes, chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(bytes == 0 || (dst_host && src_device), "chr_host_download: null pointer argument");
    std::lock_guard<std::mutex> lock(g_io.mu);
    if (int rc = host_io_ready()) return rc;
    const int nt = g_io.pool->size();
    const size_t n_chunks = (bytes + kRingSlotBytes - 1) / kRingSlotBytes;
    auto issue = [&](size_t c) -> int {
        const int s = (int)(c % kRingSlots);
        const size_t off = c * kRingSlotBytes, len = std::min(kRingSlotBytes, bytes - off);
        if (int rc = ring_acquire(s)) return rc;                        // a pending upload out of this slot
        CHR_CUDA(cudaMemcpyAsync(g_io.slot[s], (const char*)src_device + off, len, cudaMemcpyDeviceToHost, st));
        CHR_CUDA(cudaEventRecord(g_io.ev[s], st));
        g_io.busy[s] = true;
        return 0;
    };
    for (size_t c = 0; c < n_chunks && c < (size_t)kRingSlots - 1; ++c)
        if (int rc = issue(c)) return rc;
    for (size_t c = 0; c < n_chunks; ++c) {
        if (c + kRingSlots - 1 < n_chunks)
            if (int rc = issue(c + kRingSlots - 1)) return rc;          // its slot was drained in iteration c - 1
        const int s = (int)(c % kRingSlots);
        const size_t off = c * kRingSlotBytes, len = std::min(kRingSlotBytes, bytes - off);
        if (int rc = ring_acquire(s)) return rc;                        // the DMA of this chunk has landed
        const char* buf = g_io.slot[s];
        char* dst = (char*)dst_host + off;
        g_io.pool->run([&](int t) {
            const size_t per = ((len + nt - 1) / nt + 4095) & ~(size_t)4095;
            const size_t b = std::min(len, per * t), e = std::min(len, b + per);
            if (e > b) memcpy(dst + b, buf + b, e - b);
        });
    }
    return 0;
}

// count pass of the column selection: sel_indptr (HOST, int64 [n + 1]) = row pointers of the matrix restricted to the
// columns with col_map[c] >= 0 (col_map: HOST int32 [n_genes]).  Returns the selected nnz through *nnz_out.
int chr_host_csr_select_count(const int64_t* indptr, const int32_t* indices, int64_t n, int64_t n_genes, const int32_t* col_map,
                              int64_t* sel_indptr, int64_t* nnz_out) {
    using namespace chr;
    CHR_REQUIRE(indptr && indices && col_map && sel_indptr && nnz_out && n > 0 && n_genes > 0, "chr_host_csr_select_count: bad arguments");
    std::lock_guard<std::mutex> lock(g_io.mu);
    if (!g_io.pool) g_io.pool = new HostPool(host_threads());
    const int nt = g_io.pool->size();
    std::atomic<int> bad{0};
    g_io.pool->run([&](int t) {
        const int64_t per = (n + nt - 1) / nt, b = std::min(n, per * t), e = std::min(n, b + per);
        for (int64_t i = b; i < e; ++i) {
            int64_t c = 0;
            for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
                const int32_t g = indices[p];
                if ((uint32_t)g >= (uint64_t)n_genes) { bad = 1; continue; }
                c += col_map[g] >= 0;
            }
            sel_indptr[i + 1] = c;
        }
    });
    CHR_REQUIRE(!bad.load(), "chr_host_csr_select_count: column index out of range");
    sel_indptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) sel_indptr[i + 1] += sel_indptr[i];
    *nnz_out = sel_indptr[n];
    return 0;
}

// pack pass: the selected entries (new column index int32, value float32) go through the pinned ring to the device
// arrays out_indices / out_data (device, sel nnz long).  data_kind: 0 = float32, 1 = int32, 2 = float64 values.
int chr_host_csr_select_upload(const int64_t* indptr, const int32_t* indices, const void* data, int data_kind, int64_t n,
                               const int32_t* col_map, const int64_t* sel_indptr, int32_t* out_indices_dev, float* out_data_dev,
                               chr_stream_t stream) {
    using namespace chr;
    cudaStream_t st = (cudaStream_t)stream;
    CHR_REQUIRE(indptr && indices && data && col_map && sel_indptr && n > 0, "chr_host_csr_select_upload: null pointer argument");
    CHR_REQUIRE(data_kind >= 0 && data_kind <= 2, "chr_host_csr_select_upload: data_kind must be 0 (f32), 1 (i32) or 2 (f64)");
    const int64_t total = sel_indptr[n];
    CHR_REQUIRE(total == 0 || (out_indices_dev && out_data_dev), "chr_host_csr_select_upload: null output");
    std::lock_guard<std::mutex> lock(g_io.mu);
    if (int rc = host_io_ready()) return rc;
    const int nt = g_io.pool->size();
    // a chunk = a run of rows whose selected entries fit half a slot each for indices and values
    const int64_t cap = (int64_t)(kRingSlotBytes / 8);
    int s = 0;
    int64_t r0 = 0;
    while (r0 < n) {
        // rows [r0, r1): as many as fit (binary search on the row pointers); a single row longer than a slot cannot occur
        // for cap = 4 M entries unless the matrix has more columns than that
        int64_t lo = r0 + 1, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) / 2;
            if (sel_indptr[mid] - sel_indptr[r0] <= cap) lo = mid; else hi = mid - 1;
        }
        const int64_t r1 = lo;
        const int64_t base = sel_indptr[r0], cnt = sel_indptr[r1] - base;
        CHR_REQUIRE(cnt <= cap, "chr_host_csr_select_upload: a row holds more selected entries than a staging slot");
        if (int rc = ring_acquire(s)) return rc;
        int32_t* bi = (int32_t*)g_io.slot[s];
        float* bv = (float*)(g_io.slot[s] + kRingSlotBytes / 2);
        g_io.pool->run([&](int t) {
            const int64_t rows = r1 - r0, per = (rows + nt - 1) / nt, b = r0 + std::min(rows, per * t), e = std::min(r1, b + per);
            for (int64_t i = b; i < e; ++i) {
                int64_t o = sel_indptr[i] - base;
                for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
                    const int32_t c = col_map[indices[p]];
                    if (c < 0) continue;
                    bi[o] = c;
                    bv[o] = data_kind == 0 ? ((const float*)data)[p] : data_kind == 1 ? (float)((const int32_t*)data)[p] : (float)((const double*)data)[p];
                    ++o;
                }
            }
        });
        if (cnt) {
            CHR_CUDA(cudaMemcpyAsync(out_indices_dev + base, bi, (size_t)cnt * 4, cudaMemcpyHostToDevice, st));
            CHR_CUDA(cudaMemcpyAsync(out_data_dev + base, bv, (size_t)cnt * 4, cudaMemcpyHostToDevice, st));
        }
        if (int rc = ring_submit(s, nullptr, 0, st)) return rc;
        s = (s + 1) % kRingSlots;
        r0 = r1;
    }
    return 0;
}

}  // extern "C"
