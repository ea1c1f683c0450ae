// Library-level plumbing of the C ABI: version, thread-local error string, kernel timers.
#include <mutex>
#include "async.cuh"

namespace chr {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static KernelTimer g_timers[8];

KernelTimer& timer(int family) { return g_timers[family & 7]; }

void timer_begin(int family, bool on, cudaStream_t s) {
    KernelTimer& t = timer(family);
    t.armed = false;
    if (!on) return;
    {
        static std::mutex mu;                       // restarts of one fit may run on several host threads
        std::lock_guard<std::mutex> lock(mu);
        if (!t.a) {
            cudaEventCreate(&t.b);
            cudaEventCreate(&t.a);
        }
    }
    cudaEventRecord(t.a, s);
}

void timer_end(int family, bool on, cudaStream_t s) {
    if (!on) return;
    KernelTimer& t = timer(family);
    cudaEventRecord(t.b, s);
    t.armed = true;
}

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace chr

extern "C" {

int chr_version(void) { return 100; }

const char* chr_last_error(void) { return chr::g_err; }

float chr_last_kernel_ms(int family) {
    chr::KernelTimer& t = chr::timer(family);
    if (!t.armed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(t.b) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) return -1.0f;
    return ms;
}

}  // extern "C"
