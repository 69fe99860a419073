// cpu_twin.cpp -- HOST-pointer twins of the device entry points.
//
// The reference's native module dispatches on the tensor's device (binding.cpp:38,59,74:
// `if (x.is_cuda()) ... else ..._cpu(...)`); these are the `else` branches, so that the
// Python surface behaves identically for CPU tensors (BASELINE config 1 runs on CPU).
// They are NOT a fallback for the CUDA path: torchfx_b200/_ops.py only calls them for CPU
// tensors and raises if a CUDA tensor cannot be served by the CUDA kernels.
//
// Arithmetic: float64 Direct Form 1 per section, fused over sections, OpenMP over
// channels -- what the reference computes on CPU (cpu/iir_cpu.cpp:64-159).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <complex>
#include <vector>

#include "common.cuh"
#include "sos_plan.h"

namespace tfx {
namespace {

template <typename IO>
int sos_cascade_cpu(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const double *sos_host, int K,
                    double *state_x, double *state_y) {
    TFX_REQUIRE(C >= 0 && T >= 0, "sos cascade (cpu): negative shape");
    TFX_REQUIRE((state_x == nullptr) == (state_y == nullptr), "sos cascade (cpu): state_x and state_y must both be given or both NULL");
    auto plan = get_sos_plan(sos_host, K);
    if (!plan) return TFX_EINVAL;
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr, "sos cascade (cpu): NULL signal pointer");
    TFX_REQUIRE(ldx >= T && ldy >= T, "sos cascade (cpu): row stride smaller than T");
    const SosSection *sec = plan->sec.data();
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        // h[k] = {x[n-1], x[n-2], y[n-1], y[n-2]} of section k
        std::vector<double> h(static_cast<size_t>(4) * K, 0.0);
        if (state_x != nullptr) {
            for (int k = 0; k < K; ++k) {
                const int64_t o = (static_cast<int64_t>(k) * C + c) * 2;
                h[4 * k + 0] = state_x[o];
                h[4 * k + 1] = state_x[o + 1];
                h[4 * k + 2] = state_y[o];
                h[4 * k + 3] = state_y[o + 1];
            }
        }
        const IO *xr = x + c * ldx;
        IO *yr = y + c * ldy;
        for (int64_t n = 0; n < T; ++n) {
            double v = static_cast<double>(xr[n]);
            for (int k = 0; k < K; ++k) {
                double *hk = &h[4 * k];
                const SosSection &s = sec[k];
                const double out = s.b0 * v + s.b1 * hk[0] + s.b2 * hk[1] - s.a1 * hk[2] - s.a2 * hk[3];
                hk[1] = hk[0];
                hk[0] = v;
                hk[3] = hk[2];
                hk[2] = out;
                v = out;
            }
            yr[n] = static_cast<IO>(v);
        }
        if (state_x != nullptr) {
            for (int k = 0; k < K; ++k) {
                const int64_t o = (static_cast<int64_t>(k) * C + c) * 2;
                state_x[o] = h[4 * k + 0];
                state_x[o + 1] = h[4 * k + 1];
                state_y[o] = h[4 * k + 2];
                state_y[o + 1] = h[4 * k + 3];
            }
        }
    }
    return TFX_OK;
}

template <typename IO>
int delay_cpu(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay, double decay,
              double mix) {
    TFX_REQUIRE(C >= 0 && T >= 0 && delay >= 0, "delay line (cpu): negative argument");
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && x != y, "delay line (cpu): NULL or aliased buffers");
    const IO coeff = static_cast<IO>(mix * decay);
    const int64_t D = T <= delay ? T : delay;
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        const IO *xr = x + c * ldx;
        IO *yr = y + c * ldy;
        for (int64_t n = 0; n < D; ++n) yr[n] = xr[n];
        for (int64_t n = D; n < T; ++n) yr[n] = xr[n] + coeff * xr[n - D];
    }
    return TFX_OK;
}

// Causal FIR, zero history.  float32 input accumulates in double (one rounding at the end).
// Short impulse responses: direct form.  Long ones (K > kFirCpuDirectTaps): overlap-save with an in-house double-precision
// radix-2 FFT, block N >= 4 K, parallel over (channel, block) -- the reference's CPU path is overlap-save too
// (filter/_fftconv.py:107-141 on torch.fft), a direct form would be O(K T).
constexpr int64_t kFirCpuDirectTaps = 64;

using cplx = std::complex<double>;

void fft_pow2(cplx *a, int n, const cplx *w /* n/2 roots exp(-2 pi i k / n) */, bool inverse) {
    for (int i = 1, j = 0; i < n; ++i) {  // bit reversal
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                const cplx t = (inverse ? std::conj(w[k * step]) : w[k * step]) * a[i + k + half];
                a[i + k + half] = a[i + k] - t;
                a[i + k] += t;
            }
    }
}

template <typename IO>
void fir_cpu_ols(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const IO *taps, int64_t K) {
    int n = 1024;
    while (n < 4 * K && n < (1 << 22)) n <<= 1;
    while (n < 2 * K) n <<= 1;  // (only beyond 2^20 taps)
    const int64_t hop = n - K + 1;
    const int64_t nblk = (T + hop - 1) / hop;
    std::vector<cplx> w(n / 2), H(n);
    for (int k = 0; k < n / 2; ++k) {
        const double ang = -2.0 * M_PI * k / n;
        w[k] = cplx(std::cos(ang), std::sin(ang));
    }
    for (int64_t i = 0; i < n; ++i) H[i] = i < K ? cplx(static_cast<double>(taps[i]), 0.0) : cplx(0.0, 0.0);
    fft_pow2(H.data(), n, w.data(), false);
    const double scale = 1.0 / n;
#pragma omp parallel
    {
        std::vector<cplx> buf(n);
#pragma omp for collapse(2) schedule(dynamic)
        for (int64_t c = 0; c < C; ++c)
            for (int64_t b = 0; b < nblk; ++b) {
                const IO *xr = x + c * ldx;
                IO *yr = y + c * ldy;
                const int64_t s = b * hop;  // first output of the block; input window starts K - 1 samples earlier
                for (int64_t i = 0; i < n; ++i) {
                    const int64_t m = s - (K - 1) + i;
                    buf[i] = (m >= 0 && m < T) ? cplx(static_cast<double>(xr[m]), 0.0) : cplx(0.0, 0.0);
                }
                fft_pow2(buf.data(), n, w.data(), false);
                for (int64_t i = 0; i < n; ++i) buf[i] *= H[i];
                fft_pow2(buf.data(), n, w.data(), true);
                const int64_t cnt = std::min<int64_t>(hop, T - s);
                for (int64_t j = 0; j < cnt; ++j) yr[s + j] = static_cast<IO>(buf[K - 1 + j].real() * scale);
            }
    }
}

template <typename IO>
int fir_cpu(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const IO *taps, int64_t K) {
    TFX_REQUIRE(C >= 0 && T >= 0 && K >= 1, "fir (cpu): bad shape");
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && taps != nullptr && x != y, "fir (cpu): NULL or aliased buffers");
    if (K > kFirCpuDirectTaps && T > 2 * K) {
        fir_cpu_ols<IO>(x, y, C, T, ldx, ldy, taps, K);
        return TFX_OK;
    }
    std::vector<double> b(taps, taps + K);
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        const IO *xr = x + c * ldx;
        IO *yr = y + c * ldy;
        for (int64_t n = 0; n < T; ++n) {
            const int64_t jmax = std::min<int64_t>(n, K - 1);
            double acc = 0.0;
            for (int64_t j = 0; j <= jmax; ++j) acc += b[j] * static_cast<double>(xr[n - j]);
            yr[n] = static_cast<IO>(acc);
        }
    }
    return TFX_OK;
}

}  // namespace
}  // namespace tfx

extern "C" {
int tfx_sos_cascade_cpu_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                            const double *sos_host, int K, double *state_x, double *state_y) {
    return tfx::sos_cascade_cpu<float>(x, y, C, T, ldx, ldy, sos_host, K, state_x, state_y);
}
int tfx_sos_cascade_cpu_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                            const double *sos_host, int K, double *state_x, double *state_y) {
    return tfx::sos_cascade_cpu<double>(x, y, C, T, ldx, ldy, sos_host, K, state_x, state_y);
}
int tfx_delay_line_cpu_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay,
                           double decay, double mix) {
    return tfx::delay_cpu<float>(x, y, C, T, ldx, ldy, delay, decay, mix);
}
int tfx_delay_line_cpu_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t delay,
                           double decay, double mix) {
    return tfx::delay_cpu<double>(x, y, C, T, ldx, ldy, delay, decay, mix);
}
int tfx_fir_cpu_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const float *taps_host,
                    int64_t K) {
    return tfx::fir_cpu<float>(x, y, C, T, ldx, ldy, taps_host, K);
}
int tfx_fir_cpu_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                    const double *taps_host, int64_t K) {
    return tfx::fir_cpu<double>(x, y, C, T, ldx, ldy, taps_host, K);
}
}
