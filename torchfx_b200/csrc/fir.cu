// fir.cu -- causal FIR filtering, y[c,n] = sum_j taps[j] * x[c,n-j], zero history.
//
// Replaces FIR.forward (reference filter/fir.py:526-579) and its overlap-save helper
// fft_conv1d (filter/_fftconv.py:107-141), which the reference evaluates with torch.fft
// (cuFFT) on materialised frames: pad, as_strided unfold (1.25x the signal), batched rfft,
// complex multiply, batched irfft, slice -- about five HBM passes, and a single FFT block
// of int(5*K) samples (327 680 points for a 65 536-tap reverb IR).
//
// Two algorithms, both written here (no cuFFT):
//
//  DIRECT  (K <= kDirectMaxTaps): shared-memory tile of x plus the taps, 8 consecutive
//          outputs per thread with a register sliding window (1 new sample + 1 tap per 8 FMA);
//          a float64 twin for float64 signals (tfx_fir_f64).
//
//  OLS     uniformly-partitioned overlap-save as ONE persistent kernel: fir_ols16k.cu
//          (16384-point transforms in shared memory, spectra ring resident in L2).
//          Round 1's three-kernels-per-slab version (4096-point transforms, spectra and
//          products through HBM: 39.6 B/sample of DRAM traffic) lost every A/B against it
//          and was removed; its record is profiles/r1_fir.md.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "fir_ols16k.h"

namespace tfx {
namespace {

constexpr int kDirectMaxTaps = 1024;  // smem-limited; AUTO switches to OLS far earlier
constexpr int kAutoDirectTaps = 96;

// ------------------------------------------------------------------------------------------
// DIRECT
// ------------------------------------------------------------------------------------------
constexpr int kDirTile = 2048;  // outputs per CTA (256 threads x 8)

// Shared-memory index of sample i of the tile: one pad word per 32, so that the threads' windows -- 8 samples apart, i.e.
// 8 words apart: an 8-way bank conflict on every load of the main loop without the pad (ncu: mio_throttle) -- fall into
// 32 different banks, and the cooperative fill (consecutive i) stays conflict-free.
__device__ __forceinline__ int dpad(int i) { return i + (i >> 5); }

__global__ void __launch_bounds__(256) fir_direct_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t C,
                                                         int64_t T, int64_t ldx, int64_t ldy,
                                                         const float *__restrict__ taps, int K, int64_t tiles_per_row, int vec_ok) {
    extern __shared__ float sm[];
    float *bs = sm;                  // K taps
    float *xs = sm + ((K + 3) & ~3);  // K-1 history + kDirTile samples, padded (dpad)
    const int64_t c = blockIdx.x / tiles_per_row;
    const int64_t n0 = (blockIdx.x - c * tiles_per_row) * kDirTile;
    const float *xr = x + c * ldx;
    for (int i = threadIdx.x; i < K; i += 256) bs[i] = taps[i];
    const int span = kDirTile + K - 1;
    for (int i = threadIdx.x; i < span; i += 256) {
        const int64_t n = n0 - (K - 1) + i;
        xs[dpad(i)] = (n >= 0 && n < T) ? xr[n] : 0.f;
    }
    __syncthreads();
    // outputs n0 + 8t + r, r = 0..7:  y = sum_j b[j] * xs[(K-1) + 8t + r - j]
    const int t8 = threadIdx.x * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float w[8];  // w[(m) & 7] holds xs[(K-1) + t8 + m - j] for the current j, m = 0..7
#pragma unroll
    for (int m = 0; m < 8; ++m) w[m] = xs[dpad((K - 1) + t8 + m)];
    const int inew = (K - 1) + t8 - 1;  // element entering the window after tap j: sample inew - j
    int j = 0;
    for (; j + 8 <= K; j += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float b = bs[j + u];
            // window for tap j+u: output r uses xs[.. + r - (j+u)] = w[(r - u) & 7]
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(b, w[(r - u) & 7], acc[r]);
            const int i = inew - (j + u);
            w[(7 - u) & 7] = i >= 0 ? xs[dpad(i)] : 0.f;  // slot of r = 7 is free; it becomes r = 0 of the next tap (i < 0 only past the last tap)
        }
    }
    for (; j < K; ++j) {  // remainder (K % 8 taps), plain indexing
        const float b = bs[j];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(b, xs[dpad((K - 1) + t8 + r - j)], acc[r]);
    }
    float *yr = y + c * ldy;
    const int64_t nb = n0 + t8;
    if (vec_ok && nb + 8 <= T) {  // two 16-byte streaming stores instead of eight 4-byte ones
        st_stream16(yr + nb, make_float4(acc[0], acc[1], acc[2], acc[3]));
        st_stream16(yr + nb + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t n = nb + r;
            if (n < T) yr[n] = acc[r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// float64 signals: direct form in float64 for every tap count (the reference evaluates FIR in the
// input dtype, filter/fir.py:529-531).  A CTA owns 1024 outputs of one channel and walks the taps in
// chunks of 512 through shared memory; 4 consecutive outputs per thread with a register window.
// O(K T): float64 audio on the device is the rare case, the float32 overlap-save path is the fast one.
// ------------------------------------------------------------------------------------------
constexpr int kD64Tile = 1024, kD64Chunk = 512;
// one pad double per 16: the threads' windows are 4 doubles apart, which without the pad is a 4-way conflict per half-warp
__device__ __forceinline__ int dpad64(int i) { return i + (i >> 4); }
__global__ void __launch_bounds__(256) fir_direct_f64_kernel(const double *__restrict__ x, double *__restrict__ y, int64_t C, int64_t T,
                                                             int64_t ldx, int64_t ldy, const double *__restrict__ taps, int64_t K,
                                                             int64_t tiles_per_row) {
    __shared__ double bs[kD64Chunk];
    __shared__ double xs[kD64Tile + kD64Chunk + (kD64Tile + kD64Chunk) / 16 + 1];  // padded: one word per 16 (see dpad64)
    const int64_t c = blockIdx.x / tiles_per_row;
    const int64_t n0 = (blockIdx.x - c * tiles_per_row) * kD64Tile;
    const double *xr = x + c * ldx;
    const int t4 = threadIdx.x * 4;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j0 = 0; j0 < K; j0 += kD64Chunk) {
        const int kc = static_cast<int>(min(static_cast<int64_t>(kD64Chunk), K - j0));
        const int kc4 = (kc + 3) & ~3;  // taps of this chunk, zero-padded to a multiple of 4 (a 64-tap filter used to run all 512)
        __syncthreads();
        for (int i = threadIdx.x; i < kD64Chunk; i += 256) bs[i] = i < kc ? taps[j0 + i] : 0.0;
        // sample i of the tile = x[n0 - j0 - (kD64Chunk - 1) + i]: the samples taps j0 .. j0 + 511 pair with outputs n0 .. n0 + 1023
        for (int i = threadIdx.x; i < kD64Tile + kD64Chunk; i += 256) {
            const int64_t n = n0 - j0 - (kD64Chunk - 1) + i;
            xs[dpad64(i)] = (n >= 0 && n < T) ? xr[n] : 0.0;
        }
        __syncthreads();
        // output n0 + t4 + r, tap j0 + u: sample (kD64Chunk - 1) + t4 + r - u; a register window slides over the taps
        const int base = (kD64Chunk - 1) + t4;
        double w[4];  // w[(r - u) & 3] = sample base + r - u
#pragma unroll
        for (int r = 0; r < 4; ++r) w[r] = xs[dpad64(base + r)];
        for (int u = 0; u < kc4; u += 4) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const double b = bs[u + uu];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = fma(b, w[(r - uu) & 3], acc[r]);
                const int i = base - 1 - (u + uu);
                w[(3 - uu) & 3] = i >= 0 ? xs[dpad64(i)] : 0.0;  // the slot of r = 3 becomes r = 0 of the next tap
            }
        }
    }
    double *yr = y + c * ldy;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t n = n0 + t4 + r;
        if (n < T) yr[n] = acc[r];
    }
}

// TFX_FIR_AUTO.  The single-partition overlap-save path costs the same for every K <= 1024 (256 ch x 60 s: 4.2 ms), the direct
// form grows with K (1.3 / 2.3 / 2.6 / 3.4 / 4.3 / 6.0 ms at 8 / 16 / 32 / 48 / 64 / 96 taps) but has no fixed cost on small
// inputs (8 ch x 10 s: 0.05-0.09 ms against 0.10): direct up to 56 taps, and up to 96 taps below 16 M samples
// (tools/fir_direct_vs_ols.py).
constexpr int kAutoDirectTapsLarge = 56;
constexpr int64_t kAutoLargeSamples = int64_t(1) << 24;
int pick_algo(int algo, int64_t K, int64_t C, int64_t T) {
    if (algo == TFX_FIR_DIRECT || algo == TFX_FIR_OLS) return algo;
    if (K <= kAutoDirectTapsLarge) return TFX_FIR_DIRECT;
    if (K <= kAutoDirectTaps && C * T < kAutoLargeSamples) return TFX_FIR_DIRECT;
    return TFX_FIR_OLS;
}

}  // namespace
}  // namespace tfx

extern "C" {

size_t tfx_fir_workspace_bytes(int64_t C, int64_t T, int64_t K, int algo) {
    if (C <= 0 || T <= 0 || K <= 0) return 0;
    if (tfx::pick_algo(algo, K, C, T) == TFX_FIR_DIRECT && K <= tfx::kDirectMaxTaps) return 0;
    return tfx::fir_ols16k_workspace_bytes(K);
}

int tfx_fir_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const float *taps, int64_t K,
                int algo, void *workspace, size_t workspace_bytes, void *stream_v) {
    using namespace tfx;
    TFX_REQUIRE(C >= 0 && T >= 0, "fir: negative shape");
    TFX_REQUIRE(K >= 1, "fir: need at least one tap (K=%lld)", (long long)K);
    TFX_REQUIRE(algo == TFX_FIR_AUTO || algo == TFX_FIR_DIRECT || algo == TFX_FIR_OLS, "fir: bad algo %d", algo);
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && taps != nullptr && x != y, "fir: NULL or aliased buffers (not in place)");
    TFX_REQUIRE(ldx >= T && ldy >= T, "fir: row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    int use = pick_algo(algo, K, C, T);
    if (use == TFX_FIR_DIRECT && K > kDirectMaxTaps) use = TFX_FIR_OLS;

    if (use == TFX_FIR_DIRECT) {
        const int64_t tiles = (T + kDirTile - 1) / kDirTile;
        const int64_t span = kDirTile + K - 1;
        const size_t smem = sizeof(float) * (((K + 3) & ~3) + span + span / 32 + 1);
        const int vec_ok = (reinterpret_cast<uintptr_t>(y) % 16 == 0 && ldy % 4 == 0) ? 1 : 0;
        TFX_REQUIRE(C * tiles < (int64_t(1) << 31), "fir: too many tiles for one launch");
        fir_direct_kernel<<<static_cast<unsigned>(C * tiles), 256, smem, stream>>>(x, y, C, T, ldx, ldy, taps, static_cast<int>(K),
                                                                                 tiles, vec_ok);
        TFX_CHECK_LAUNCH("fir_direct_kernel");
        return TFX_OK;
    }

    return launch_fir_ols16k(x, y, C, T, ldx, ldy, taps, nullptr, K, workspace, workspace_bytes, stream);
}

size_t tfx_fir_plan_bytes(int64_t K) { return K <= 0 ? 0 : tfx::fir_ols16k_plan_bytes(K); }

int tfx_fir_plan_init(const float *taps, int64_t K, void *plan, size_t plan_bytes, void *stream_v) {
    using namespace tfx;
    TFX_REQUIRE(K >= 1, "fir plan: need at least one tap (K=%lld)", (long long)K);
    TFX_REQUIRE(taps != nullptr, "fir plan: NULL taps");
    int rc = require_device();
    if (rc != TFX_OK) return rc;
    return fir_ols16k_plan_init(taps, K, plan, plan_bytes, static_cast<cudaStream_t>(stream_v));
}

int tfx_fir_f32_planned(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const void *plan, int64_t K,
                        void *workspace, size_t workspace_bytes, void *stream_v) {
    using namespace tfx;
    TFX_REQUIRE(C >= 0 && T >= 0, "fir: negative shape");
    TFX_REQUIRE(K >= 1, "fir: need at least one tap (K=%lld)", (long long)K);
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && plan != nullptr && x != y, "fir: NULL or aliased buffers (not in place)");
    TFX_REQUIRE(ldx >= T && ldy >= T, "fir: row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;
    return launch_fir_ols16k(x, y, C, T, ldx, ldy, nullptr, plan, K, workspace, workspace_bytes, static_cast<cudaStream_t>(stream_v));
}

int tfx_fir_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const double *taps, int64_t K,
                void *stream_v) {
    using namespace tfx;
    TFX_REQUIRE(C >= 0 && T >= 0, "fir: negative shape");
    TFX_REQUIRE(K >= 1, "fir: need at least one tap (K=%lld)", (long long)K);
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && taps != nullptr && x != y, "fir: NULL or aliased buffers (not in place)");
    TFX_REQUIRE(ldx >= T && ldy >= T, "fir: row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;
    const int64_t tiles = (T + kD64Tile - 1) / kD64Tile;
    TFX_REQUIRE(C * tiles < (int64_t(1) << 31), "fir: too many tiles for one launch");
    fir_direct_f64_kernel<<<static_cast<unsigned>(C * tiles), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(x, y, C, T, ldx, ldy, taps, K, tiles);
    TFX_CHECK_LAUNCH("fir_direct_f64_kernel");
    return TFX_OK;
}

}  // extern "C"
