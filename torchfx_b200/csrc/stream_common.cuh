// stream_common.cuh -- pieces shared by the streaming recurrence kernels
// (sos_cascade.cu, filterbank.cu).
#pragma once

#include "common.cuh"

namespace tfx {

constexpr int kRowBytes = 256;  // one chunk of one stream in shared memory
constexpr int kPitch = 272;     // row pitch (256 + 16): per-lane 128-bit row accesses are bank-conflict free

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

__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

}  // namespace tfx
