// common.cuh -- error convention, launch accounting and small device helpers shared by
// every translation unit of libtorchfx_b200.so.  No torch headers anywhere in csrc/.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "torchfx_b200.h"

namespace tfx {

// Thread-local error message (returned by tfx_last_error()).
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
// Adds n to the process-wide kernel-launch counter (tfx_kernel_launches()).
void count_launch(int n = 1);
// TFX_OK if a CUDA device is usable, else TFX_ENODEVICE (with message).  Never computes
// on the CPU instead: callers propagate the error.
int require_device();
// SM count of the current device (cached per device id).
int sm_count();
// Index of the current device, clamped to [0, 64): slot for per-device one-time setup flags.
int device_slot();

#define TFX_CUDA_TRY(expr)                                                        \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) return ::tfx::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define TFX_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ::tfx::set_error(__VA_ARGS__);    \
            return TFX_EINVAL;                \
        }                                     \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device): do it once per
// device for the kernel instantiation this macro is expanded in (racing threads set the same value).
#define TFX_ENSURE_SMEM(kern, bytes)                                                                          \
    do {                                                                                                     \
        static bool done__[64] = {};                                                                         \
        const int slot__ = ::tfx::device_slot();                                                             \
        if (!done__[slot__]) {                                                                               \
            TFX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));  \
            done__[slot__] = true;                                                                           \
        }                                                                                                    \
    } while (0)

// Checks the launch that was just issued.
#define TFX_CHECK_LAUNCH(name)                                                         \
    do {                                                                               \
        ::tfx::count_launch();                                                         \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) return ::tfx::cuda_fail(e__, name, __FILE__, __LINE__); \
    } while (0)

#ifdef __CUDACC__
// ---- cp.async (LDGSTS) helpers --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16) {
        // .cg: bypass L1 -- every byte of the stream is read exactly once.
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                     : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                     "n"(BYTES)
                     : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// Streaming (evict-first) 16-byte global store: output is written once and never re-read.
__device__ __forceinline__ void st_stream16(void *gptr, const float4 &v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(gptr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream16(void *gptr, const double2 &v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(gptr), "d"(v.x), "d"(v.y) : "memory");
}
#endif  // __CUDACC__

}  // namespace tfx
