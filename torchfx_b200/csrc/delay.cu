// delay.cu -- feed-forward delay line, y[n] = x[n] + (mix*decay) * x[n - D]  (n >= D).
//
// Replaces (reference) cuda/delay_forward.cu:15-123 / cpu/delay_cpu.cpp:17-85.  A pure
// streaming kernel: 8 B/sample algorithmic.  64-bit indexing throughout (the reference's
// `int idx = channel*T + n`, delay_forward.cu:29, overflows at 2^31 elements).
//
// A CTA owns one tile of 4096 (float32) / 2048 (float64) consecutive samples of ONE row and CTAs walk along a
// row in launch order, so the delayed tap x[n - D] of a tile is what the CTAs just before it streamed in: it is
// served by L2.  (The first version handed the vectors out grid-strided across all rows: the tap came back from
// DRAM -- ncu: 19.3 GB read for 11.8 GB of input -- and every access was a scalar 4-byte one: 8.6 ms for
// 1024 ch x 60 s; this one moves 16-byte vectors for x and y and reads the tap with the alignment it has.)
#include "common.cuh"

namespace tfx {
namespace {

constexpr int kDelayVecsPerThread = 4;

template <typename IO>
__global__ void __launch_bounds__(256) delay_kernel(const IO *__restrict__ x, IO *__restrict__ y, int64_t C, int64_t T,
                                                    int64_t ldx, int64_t ldy, int64_t D, IO coeff, int64_t tiles_per_row, int vec_ok) {
    constexpr int VEC = 16 / sizeof(IO);
    constexpr int TILE = 256 * kDelayVecsPerThread * VEC;
    const int64_t c = blockIdx.x / tiles_per_row;
    const int64_t t0 = (blockIdx.x - c * tiles_per_row) * TILE;
    const IO *xr = x + c * ldx;
    IO *yr = y + c * ldy;
#pragma unroll
    for (int v = 0; v < kDelayVecsPerThread; ++v) {
        const int64_t n0 = t0 + (static_cast<int64_t>(v) * 256 + threadIdx.x) * VEC;
        if (n0 >= T) break;
        if (vec_ok && n0 + VEC <= T) {
            IO xn[VEC], out[VEC];
            if constexpr (sizeof(IO) == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(xr + n0);
                xn[0] = a.x, xn[1] = a.y, xn[2] = a.z, xn[3] = a.w;
            } else {
                const double2 a = *reinterpret_cast<const double2 *>(xr + n0);
                xn[0] = a.x, xn[1] = a.y;
            }
            if (n0 >= D) {
                const IO *xd = xr + (n0 - D);
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = fma(coeff, __ldg(xd + e), xn[e]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = (n0 + e >= D) ? fma(coeff, __ldg(xr + (n0 + e - D)), xn[e]) : xn[e];
            }
            if constexpr (sizeof(IO) == 4)
                st_stream16(yr + n0, make_float4(out[0], out[1], out[2], out[3]));
            else
                st_stream16(yr + n0, make_double2(out[0], out[1]));
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int64_t n = n0 + e;
                if (n < T) {
                    const IO xn = xr[n];
                    yr[n] = n >= D ? fma(coeff, xr[n - D], xn) : xn;
                }
            }
        }
    }
}

template <typename IO>
int delay_device(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay, double decay,
                 double mix, void *stream_v) {
    TFX_REQUIRE(C >= 0 && T >= 0 && delay >= 0, "delay line: negative argument");
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && x != y, "delay line: NULL or aliased buffers (not in place)");
    TFX_REQUIRE(ldx >= T && ldy >= T, "delay line: row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    // T <= delay: the reference returns the input unchanged (delay_forward.cu:88-90).
    const int64_t D = T <= delay ? T : delay;
    constexpr int64_t tile = 256 * kDelayVecsPerThread * (16 / sizeof(IO));
    const int64_t tiles_per_row = (T + tile - 1) / tile;
    TFX_REQUIRE(C * tiles_per_row < (int64_t(1) << 31), "delay line: too many tiles for one launch");
    const int vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                       ((ldx * sizeof(IO)) % 16 == 0) && ((ldy * sizeof(IO)) % 16 == 0);
    delay_kernel<IO><<<static_cast<unsigned>(C * tiles_per_row), 256, 0, stream>>>(x, y, C, T, ldx, ldy, D, static_cast<IO>(mix * decay),
                                                                                 tiles_per_row, vec_ok);
    TFX_CHECK_LAUNCH("delay_kernel");
    return TFX_OK;
}

}  // namespace
}  // namespace tfx

extern "C" {
int tfx_delay_line_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay,
                       double decay, double mix, void *stream) {
    return tfx::delay_device<float>(x, y, C, T, ldx, ldy, delay, decay, mix, stream);
}
int tfx_delay_line_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay,
                       double decay, double mix, void *stream) {
    return tfx::delay_device<double>(x, y, C, T, ldx, ldy, delay, decay, mix, stream);
}
}
