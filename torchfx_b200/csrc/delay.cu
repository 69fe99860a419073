// delay.cu -- feed-forward delay line, y[n] = x[n] + (mix*decay) * x[n - D]  (n >= D).
//
// Replaces (reference) cuda/delay_forward.cu:15-123 / cpu/delay_cpu.cpp:17-85.  A pure
// streaming kernel: 8 B/sample (the delayed tap is re-read from L2, D samples behind the
// front).  64-bit indexing throughout (the reference's `int idx = channel*T + n`,
// delay_forward.cu:29, overflows at 2^31 elements); 4 samples per thread.
#include "common.cuh"

namespace tfx {
namespace {

template <typename IO>
__global__ void __launch_bounds__(256) delay_kernel(const IO *__restrict__ x, IO *__restrict__ y, int64_t C, int64_t T,
                                                    int64_t ldx, int64_t ldy, int64_t D, IO coeff, int64_t nvec_row) {
    const int64_t total = C * nvec_row;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t c = i / nvec_row;
        const int64_t n0 = (i - c * nvec_row) * 4;
        const IO *xr = x + c * ldx;
        IO *yr = y + c * ldy;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t n = n0 + e;
            if (n < T) {
                const IO xn = xr[n];
                yr[n] = n >= D ? fma(coeff, xr[n - D], xn) : xn;
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
    const int64_t nvec_row = (T + 3) / 4;
    const int64_t total = C * nvec_row;
    const int64_t blocks = std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count()) * 16);
    delay_kernel<IO><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, y, C, T, ldx, ldy, D,
                                                                       static_cast<IO>(mix * decay), nvec_row);
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
