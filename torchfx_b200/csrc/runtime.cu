// runtime.cu -- library-level entry points: error state, device probing, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace tfx {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    set_error("CUDA error %d (%s) at %s [%s:%d]", static_cast<int>(e), cudaGetErrorString(e), what,
              base ? base + 1 : file, line);
    // cudaErrorNoDevice / InsufficientDriver mean "no GPU here", not a kernel bug.
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return TFX_ENODEVICE;
    return TFX_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_error("no usable CUDA device (cudaGetDeviceCount: %s); this entry point runs sm_100a kernels only and has no CPU fallback",
                  e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return TFX_ENODEVICE;
    }
    return TFX_OK;
}

int device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

int sm_count() {
    static std::mutex mu;
    static int cached[64];
    static bool have[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return 148;  // B200
    }
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            n = 148;
        }
        cached[dev] = n;
        have[dev] = true;
    }
    return cached[dev];
}

}  // namespace tfx

extern "C" {

int tfx_version(void) { return TFX_VERSION; }
const char *tfx_last_error(void) { return tfx::g_err; }
int tfx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}
uint64_t tfx_kernel_launches(void) { return tfx::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
