// stream_common.cuh -- pieces shared by the streaming recurrence kernels
// (sos_cascade.cu, filterbank.cu).
#pragma once

#include "common.cuh"

namespace tfx {

#ifndef TFX_ROWBYTES
#define TFX_ROWBYTES 256
#endif
constexpr int kRowBytes = TFX_ROWBYTES;  // one chunk of one stream in shared memory
constexpr int kPitch = kRowBytes + 16;   // row pitch: per-lane 128-bit row accesses are bank-conflict free
static_assert((kPitch / 4) % 32 == 4, "row pitch must be 4 words mod 32 for conflict-free 128-bit row accesses");
constexpr int kNvec = kRowBytes / 16;    // 16-byte vectors per row
constexpr int kSmemPerSm = 227 * 1024;

template <typename IO>
struct IoTraits;
template <>
struct IoTraits<float> {
    using Vec = float4;
    static constexpr int VEC = 4;
    static constexpr int CHUNK = kRowBytes / 4;
};
template <>
struct IoTraits<double> {
    using Vec = double2;
    static constexpr int VEC = 2;
    static constexpr int CHUNK = kRowBytes / 8;
};

#ifdef __CUDACC__
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
#endif

template <typename CT, int K>
struct SosCoef {
    CT b0[K], b1[K], b2[K], na1[K], na2[K];
};
template <int K>
struct SosCoefD {
    double b0[K], b1[K], b2[K], a1[K], a2[K];
};


#ifdef __CUDACC__
template <typename CT, int K>
__device__ __forceinline__ CT sos_step(const SosCoef<CT, K> &cf, CT (&s1)[K], CT (&s2)[K], CT v) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const CT y = fma_rn(cf.b0[k], v, s1[k]);
        s1[k] = fma_rn(cf.na1[k], y, fma_rn(cf.b1[k], v, s2[k]));
        s2[k] = fma_rn(cf.na2[k], y, cf.b2[k] * v);
        v = y;
    }
    return v;
}

template <typename CT, int K>
__device__ __forceinline__ void filter_vec(const SosCoef<CT, K> &cf, CT (&s1)[K], CT (&s2)[K], float4 &a) {
    a.x = static_cast<float>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.x)));
    a.y = static_cast<float>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.y)));
    a.z = static_cast<float>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.z)));
    a.w = static_cast<float>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.w)));
}
template <typename CT, int K>
__device__ __forceinline__ void filter_vec(const SosCoef<CT, K> &cf, CT (&s1)[K], CT (&s2)[K], double2 &a) {
    a.x = static_cast<double>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.x)));
    a.y = static_cast<double>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(a.y)));
}

#endif

}  // namespace tfx
