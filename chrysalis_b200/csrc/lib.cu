// Library-level plumbing of the C ABI: version, thread-local error string, kernel timers.
#include <cstdlib>
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

// one event pair per (device, family): events belong to the device that was current when they were created, so a
// stream of another device gets its own pair.  Creation and recording are serialised; a failed record disarms the
// timer and is cleared, so it can never surface as the error of the kernel launch that follows.
constexpr int kTimerDevices = 16, kTimerFamilies = 8;
static KernelTimer g_timers[kTimerDevices][kTimerFamilies];
static std::mutex g_timer_mu;

static int timer_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev & (kTimerDevices - 1);
}

KernelTimer& timer(int family) { return g_timers[timer_device()][family & (kTimerFamilies - 1)]; }

void timer_begin(int family, bool on, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_timer_mu);
    KernelTimer& t = timer(family);
    t.armed = false;
    if (!on) return;
    if (!t.a && (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess)) {
        cudaGetLastError();
        t.a = t.b = nullptr;
        return;
    }
    if (cudaEventRecord(t.a, s) != cudaSuccess) { cudaGetLastError(); return; }
    t.started = true;
}

void timer_end(int family, bool on, cudaStream_t s) {
    if (!on) return;
    std::lock_guard<std::mutex> lock(g_timer_mu);
    KernelTimer& t = timer(family);
    if (!t.started) return;
    t.started = false;
    if (cudaEventRecord(t.b, s) != cudaSuccess) { cudaGetLastError(); return; }
    t.armed = true;
}

bool timing_env() {
    static const bool on = [] { const char* e = getenv("CHR_TIMING"); return e && e[0] == '1'; }();
    return on;
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
