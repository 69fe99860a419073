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
//          outputs per thread with a register sliding window (1 new sample + 1 tap per 8 FMA).
//
//  OLS     uniformly-partitioned overlap-save.  The impulse response is cut into P = ceil(K/B)
//          partitions of B = 2048 taps; every block of B output samples costs ONE forward and
//          ONE inverse 4096-point complex FFT done entirely in shared memory (radix-4,
//          6 passes, 256 threads), plus a frequency-domain multiply-accumulate over the P most
//          recent input spectra (the "frequency-domain delay line"):
//              Y_k = sum_p H_p . X_{k-p},   y[kB:(k+1)B] = IFFT(Y_k)[B:2B]
//          * Two real channels ride in one complex FFT (z = x_a + i x_b): the filter is real,
//            so Re/Im of the inverse transform are the two channels' outputs -- no real-FFT
//            split pass and half the transforms.
//          * Forward = decimation-in-frequency (natural in, digit-reversed out), inverse =
//            decimation-in-time (digit-reversed in, natural out); the spectra are only ever
//            multiplied point-wise, so nothing is ever re-ordered.
//          * The multiply-accumulate is a length-P complex FIR along the block index for every
//            bin: a CTA stages a [95 blocks x 64 bins] tile of spectra and the [32 x 64] taps
//            in shared memory, each thread slides a register window over 8 consecutive blocks
//            (2 shared loads per 32 FMA), so spectra are read ~1.5x instead of P x.
//          * Long signals are processed in time slabs so the spectra workspace stays bounded.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "fir_ols16k.h"
#include "tma.cuh"

namespace tfx {
namespace {

constexpr int kDirectMaxTaps = 1024;  // smem-limited; AUTO switches to OLS far earlier
constexpr int kAutoDirectTaps = 96;
constexpr int kN = 4096;              // complex FFT size
constexpr int kB = 2048;              // partition / hop
constexpr int kFftThreads = 256;
constexpr int kMacBins = 64;          // bins per MAC CTA
constexpr int kMacBlocks = 128;       // output blocks per MAC tile
constexpr int kMacPerThread = 16;     // consecutive blocks per thread (512 threads = 64 bins x 8 groups)
constexpr int kMacPc = 32;            // partitions per shared-memory chunk
constexpr int kMacRows = kMacBlocks + kMacPc - 1;  // 159 spectra rows staged per chunk
// 1: stage the MAC tiles with cp.async.bulk + mbarriers instead of LDGSTS.  Measured SLOWER on B200 (config 3:
// 12.08 ms against 8.55 ms; profiles/r1_fir.md) -- 191 512-byte bulk copies per item run into the per-SM bulk-copy
// request rate -- so it is off; kept as a build variant (tools/build_stack_variants.sh, UNIT=fir).
#ifndef TFX_MAC_BULK
#define TFX_MAC_BULK 0
#endif

// ------------------------------------------------------------------------------------------
// DIRECT
// ------------------------------------------------------------------------------------------
constexpr int kDirTile = 2048;  // outputs per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) fir_direct_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t C,
                                                         int64_t T, int64_t ldx, int64_t ldy,
                                                         const float *__restrict__ taps, int K, int64_t tiles_per_row) {
    extern __shared__ float sm[];
    float *bs = sm;                  // K taps
    float *xs = sm + ((K + 3) & ~3);  // K-1 history + kDirTile samples
    const int64_t c = blockIdx.x / tiles_per_row;
    const int64_t n0 = (blockIdx.x - c * tiles_per_row) * kDirTile;
    const float *xr = x + c * ldx;
    for (int i = threadIdx.x; i < K; i += 256) bs[i] = taps[i];
    const int span = kDirTile + K - 1;
    for (int i = threadIdx.x; i < span; i += 256) {
        const int64_t n = n0 - (K - 1) + i;
        xs[i] = (n >= 0 && n < T) ? xr[n] : 0.f;
    }
    __syncthreads();
    // outputs n0 + 8t + r, r = 0..7:  y = sum_j b[j] * xs[(K-1) + 8t + r - j]
    const int t8 = threadIdx.x * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float w[8];  // w[(m) & 7] holds xs[(K-1) + t8 + m - j] for the current j, m = 0..7
#pragma unroll
    for (int m = 0; m < 8; ++m) w[m] = xs[(K - 1) + t8 + m];
    const float *xnew = xs + (K - 1) + t8 - 1;  // element entering the window after tap j: xnew[-j]
    int j = 0;
    for (; j + 8 <= K; j += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float b = bs[j + u];
            // window for tap j+u: output r uses xs[.. + r - (j+u)] = w[(r - u) & 7]
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(b, w[(r - u) & 7], acc[r]);
            w[(7 - u) & 7] = xnew[-(j + u)];  // slot of r = 7 is free; it becomes r = 0 of the next tap
        }
    }
    for (; j < K; ++j) {  // remainder (K % 8 taps), plain indexing
        const float b = bs[j];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(b, xs[(K - 1) + t8 + r - j], acc[r]);
    }
    float *yr = y + c * ldy;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t n = n0 + t8 + r;
        if (n < T) yr[n] = acc[r];
    }
}

// ------------------------------------------------------------------------------------------
// float64 signals: direct form in float64 for every tap count (the reference evaluates FIR in the
// input dtype, filter/fir.py:529-531).  A CTA owns 1024 outputs of one channel and walks the taps in
// chunks of 512 through shared memory; 4 consecutive outputs per thread with a register window.
// O(K T): float64 audio on the device is the rare case, the float32 overlap-save path is the fast one.
// ------------------------------------------------------------------------------------------
constexpr int kD64Tile = 1024, kD64Chunk = 512;
__global__ void __launch_bounds__(256) fir_direct_f64_kernel(const double *__restrict__ x, double *__restrict__ y, int64_t C, int64_t T,
                                                             int64_t ldx, int64_t ldy, const double *__restrict__ taps, int64_t K,
                                                             int64_t tiles_per_row) {
    __shared__ double bs[kD64Chunk];
    __shared__ double xs[kD64Tile + kD64Chunk];
    const int64_t c = blockIdx.x / tiles_per_row;
    const int64_t n0 = (blockIdx.x - c * tiles_per_row) * kD64Tile;
    const double *xr = x + c * ldx;
    const int t4 = threadIdx.x * 4;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j0 = 0; j0 < K; j0 += kD64Chunk) {
        const int kc = static_cast<int>(min(static_cast<int64_t>(kD64Chunk), K - j0));
        __syncthreads();
        for (int i = threadIdx.x; i < kD64Chunk; i += 256) bs[i] = i < kc ? taps[j0 + i] : 0.0;
        // xs[i] = x[n0 - j0 - (kD64Chunk - 1) + i]: the samples taps j0 .. j0 + 511 pair with outputs n0 .. n0 + 1023
        for (int i = threadIdx.x; i < kD64Tile + kD64Chunk; i += 256) {
            const int64_t n = n0 - j0 - (kD64Chunk - 1) + i;
            xs[i] = (n >= 0 && n < T) ? xr[n] : 0.0;
        }
        __syncthreads();
        // output n0 + t4 + r, tap j0 + u: x[n0 + t4 + r - j0 - u] = xs[(kD64Chunk - 1) + t4 + r - u]
        const double *w = xs + (kD64Chunk - 1) + t4;
#pragma unroll 4
        for (int u = 0; u < kD64Chunk; ++u) {
            const double b = bs[u];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fma(b, w[r - u], acc[r]);
        }
    }
    double *yr = y + c * ldy;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t n = n0 + t4 + r;
        if (n < T) yr[n] = acc[r];
    }
}

// ------------------------------------------------------------------------------------------
// Shared-memory radix-4 FFT (N = 4096 complex, 256 threads, in place)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// The transform is three radix-16 passes (two radix-4 stages fused in registers per pass), one
// radix-16 butterfly per thread per pass: 3 shared-memory round trips instead of 6.  Logical
// index i lives at i + (i >> 4) (one pad slot per 16) so that the stride-1, stride-16 and
// stride-256 accesses of the three passes are all (nearly) bank-conflict free.
constexpr int kPadN = kN + kN / 16;
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }

// radix-4 DIF butterfly on registers; twiddles applied to outputs 1..3 (w1, w2, w3)
__device__ __forceinline__ void bfly4_dif(float2 &a, float2 &b, float2 &c, float2 &d, bool tw, float2 w1, float2 w2, float2 w3) {
    const float2 t0 = make_float2(a.x + c.x, a.y + c.y), t1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 t2 = make_float2(b.x + d.x, b.y + d.y);
    const float2 t3 = make_float2(b.y - d.y, -(b.x - d.x));  // (b - d) * (-i)
    a = make_float2(t0.x + t2.x, t0.y + t2.y);
    float2 y1 = make_float2(t1.x + t3.x, t1.y + t3.y);
    float2 y2 = make_float2(t0.x - t2.x, t0.y - t2.y);
    float2 y3 = make_float2(t1.x - t3.x, t1.y - t3.y);
    if (tw) {
        y1 = cmul(y1, w1);
        y2 = cmul(y2, w2);
        y3 = cmul(y3, w3);
    }
    b = y1;
    c = y2;
    d = y3;
}
// radix-4 DIT (inverse) butterfly on registers; conjugate twiddles applied to inputs 1..3
__device__ __forceinline__ void bfly4_dit_inv(float2 &a, float2 &b, float2 &c, float2 &d, bool tw, float2 w1, float2 w2, float2 w3) {
    if (tw) {
        b = cmulc(b, w1);
        c = cmulc(c, w2);
        d = cmulc(d, w3);
    }
    const float2 t0 = make_float2(a.x + c.x, a.y + c.y), t1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 t2 = make_float2(b.x + d.x, b.y + d.y);
    const float2 t3 = make_float2(-(b.y - d.y), b.x - d.x);  // (b - d) * (+i)
    a = make_float2(t0.x + t2.x, t0.y + t2.y);
    b = make_float2(t1.x + t3.x, t1.y + t3.y);
    c = make_float2(t0.x - t2.x, t0.y - t2.y);
    d = make_float2(t1.x - t3.x, t1.y - t3.y);
}

// W_16^k = exp(-2*pi*i*k/16).  Stage-A twiddles factor as W_{16q}^{(j + m q) r} = W_{16q}^{j r} * W_16^{m r}:
// three table loads per pass instead of twelve, the rest are these compile-time constants.
__device__ __forceinline__ float2 w16(int k) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    switch (k & 15) {
        case 0: return make_float2(1.f, 0.f);
        case 1: return make_float2(c1, -s1);
        case 2: return make_float2(h, -h);
        case 3: return make_float2(s1, -c1);
        case 4: return make_float2(0.f, -1.f);
        case 5: return make_float2(-s1, -c1);
        case 6: return make_float2(-h, -h);
        case 7: return make_float2(-c1, -s1);
        case 8: return make_float2(-1.f, 0.f);
        case 9: return make_float2(-c1, s1);
        default: return make_float2(0.f, 0.f);  // 10..15 never needed (m, r <= 3)
    }
}

// Twiddles of one thread for one pass, W = exp(-2*pi*i/N), j = the thread's position inside its
// 16q-point group:  a_r = W^{r * j * N/(16q)},  w_r = W^{r * j * N/(4q)},  r = 1..3.
// They are tabulated per (pass, j) as six consecutive float2 (48 bytes, three 16-byte loads that
// neighbouring lanes coalesce) -- the first version gathered them from one N-entry table with
// strides up to 96 bytes, and those gathers cost as many L1 cycles as the shared-memory traffic.
struct Tw6 {
    float2 a1, a2, a3, w1, w2, w3;
};
constexpr int kTwEntries = 256 + 16;  // pass q = 256 (j < 256) then pass q = 16 (j < 16)
constexpr size_t kTwBytes = sizeof(Tw6) * kTwEntries;
template <int LQ>
__device__ __forceinline__ Tw6 load_tw(const float4 *__restrict__ tab, int j) {
    Tw6 t;
    if constexpr (LQ == 0) {
        t.a1 = t.a2 = t.a3 = t.w1 = t.w2 = t.w3 = make_float2(1.f, 0.f);
    } else {
        const float4 *e = tab + 3 * ((LQ == 8 ? 0 : 256) + j);
        const float4 u0 = __ldg(e), u1 = __ldg(e + 1), u2 = __ldg(e + 2);
        t.a1 = make_float2(u0.x, u0.y);
        t.a2 = make_float2(u0.z, u0.w);
        t.a3 = make_float2(u1.x, u1.y);
        t.w1 = make_float2(u1.z, u1.w);
        t.w2 = make_float2(u2.x, u2.y);
        t.w3 = make_float2(u2.z, u2.w);
    }
    return t;
}

// One radix-16 step on 16 registers (two fused radix-4 stages).  Forward = decimation in frequency
// (stage A on stride 4q, then stage B on stride q); inverse = the mirror image with conjugate twiddles.
template <int LQ>
__device__ __forceinline__ void radix16_dif(float2 (&v)[16], const Tw6 &t) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
        bfly4_dif(v[m], v[m + 4], v[m + 8], v[m + 12], true, m ? cmul(t.a1, w16(m)) : t.a1, m ? cmul(t.a2, w16(2 * m)) : t.a2,
                  m ? cmul(t.a3, w16(3 * m)) : t.a3);
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4_dif(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3], LQ > 0, t.w1, t.w2, t.w3);
}
template <int LQ>
__device__ __forceinline__ void radix16_dit_inv(float2 (&v)[16], const Tw6 &t) {
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4_dit_inv(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3], LQ > 0, t.w1, t.w2, t.w3);
#pragma unroll
    for (int m = 0; m < 4; ++m)
        bfly4_dit_inv(v[m], v[m + 4], v[m + 8], v[m + 12], true, m ? cmul(t.a1, w16(m)) : t.a1, m ? cmul(t.a2, w16(2 * m)) : t.a2,
                      m ? cmul(t.a3, w16(3 * m)) : t.a3);
}

// The 4096-point transform is three radix-16 passes with strides q = 256, 16, 1 (forward) or 1, 16,
// 256 (inverse), 256 threads, one radix-16 per thread per pass.  Only the MIDDLE pass lives entirely
// in shared memory: the stride-256 pass reads (forward) or writes (inverse) element j + 256 m from
// thread j -- coalesced straight from / to global memory -- and the stride-1 pass hands thread t the
// 16 consecutive points 16 t + m, which are stored to global memory TRANSPOSED, at m * 256 + t
// (again coalesced).  The spectrum order is therefore "digit-reversed, then transposed"; spectra are
// only ever multiplied point-wise with spectra in the same order, so the order never has to be undone.
// Shared-memory traffic per transform: 2 writes + 2 reads of the 32 KB tile instead of 4 + 4.
__device__ __forceinline__ void fft_fwd_tail(float2 (&v)[16], float2 *s, const float4 *__restrict__ tab,
                                             float2 *__restrict__ out) {
    const int tid = threadIdx.x;
    // pass q = 256 (registers were filled by the caller with elements tid + 256 m)
    radix16_dif<8>(v, load_tw<8>(tab, tid));
#pragma unroll
    for (int m = 0; m < 16; ++m) s[pidx(tid + 256 * m)] = v[m];
    __syncthreads();
    {   // pass q = 16
        const int j = tid & 15, base = ((tid >> 4) << 8) + j;
        const Tw6 t = load_tw<4>(tab, j);
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = s[pidx(base + 16 * m)];
        radix16_dif<4>(v, t);
#pragma unroll
        for (int m = 0; m < 16; ++m) s[pidx(base + 16 * m)] = v[m];
    }
    __syncthreads();
    // pass q = 1
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = s[pidx(16 * tid + m)];
    radix16_dif<0>(v, load_tw<0>(tab, 0));
#pragma unroll
    for (int m = 0; m < 16; ++m) out[m * 256 + tid] = v[m];
}

// inverse (unscaled): `in` in the forward transform's output order, result element tid + 256 m in v[m]
__device__ __forceinline__ void fft_inv(float2 (&v)[16], float2 *s, const float4 *__restrict__ tab,
                                        const float2 *__restrict__ in) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = __ldcs(&in[m * 256 + tid]);
    radix16_dit_inv<0>(v, load_tw<0>(tab, 0));
#pragma unroll
    for (int m = 0; m < 16; ++m) s[pidx(16 * tid + m)] = v[m];
    __syncthreads();
    {
        const int j = tid & 15, base = ((tid >> 4) << 8) + j;
        const Tw6 t = load_tw<4>(tab, j);
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = s[pidx(base + 16 * m)];
        radix16_dit_inv<4>(v, t);
#pragma unroll
        for (int m = 0; m < 16; ++m) s[pidx(base + 16 * m)] = v[m];
    }
    __syncthreads();
    const Tw6 t = load_tw<8>(tab, tid);
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = s[pidx(tid + 256 * m)];
    radix16_dit_inv<8>(v, t);
}

__global__ void fir_twiddle_kernel(float2 *tab) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // entry: pass q = 256 for e < 256, else q = 16
    if (e < kTwEntries) {
        const int ta = e < 256 ? e : (e - 256) << 4;  // j * N / (16 q)
        for (int r = 0; r < 6; ++r) {
            const int k = (r < 3 ? (r + 1) : 4 * (r - 2)) * ta;
            double sn, cs;
            sincospi(-2.0 * k / kN, &sn, &cs);
            tab[6 * e + r] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
        }
    }
}

// H[p] = FFT(taps[pB:(p+1)B] zero-padded to N), one CTA per partition
__global__ void __launch_bounds__(kFftThreads) fir_taps_fft_kernel(const float *__restrict__ taps, int64_t K, float2 *__restrict__ H,
                                                                  const float4 *__restrict__ tab) {
    __shared__ float2 s[kPadN];
    const int64_t p = blockIdx.x;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = threadIdx.x + 256 * m;
        const int64_t j = p * kB + i;
        v[m] = make_float2((i < kB && j < K) ? __ldg(&taps[j]) : 0.f, 0.f);
    }
    fft_fwd_tail(v, s, tab, H + p * kN);
}

// Z is a ring of `nrows` spectra per pair: slab row r (block k = k_first + r) lives at ring slot
// (ring0 + r) mod nrows, so the P-1 history spectra a slab needs are simply still there from the
// previous slab.  Rows [row0, row0 + gridDim.x) are computed here.
__device__ __forceinline__ int64_t ring_slot(int64_t ring0, int64_t row, int64_t nrows) {
    const int64_t r = ring0 + row;
    return r >= nrows ? r - nrows : r;
}
// Z[pair][row] = FFT(x_a[(k-1)B : (k+1)B] + i x_b[...]),  k = k_first + row (k < 0 -> zeros)
__global__ void __launch_bounds__(kFftThreads) fir_fwd_kernel(const float *__restrict__ x, int64_t C, int64_t T, int64_t ldx,
                                                             int64_t k_first, int64_t nrows, int64_t row0, int64_t ring0,
                                                             float2 *__restrict__ Z, const float4 *__restrict__ tab) {
    __shared__ float2 s[kPadN];
    const int64_t row = row0 + blockIdx.x;
    const int64_t pair = blockIdx.y;
    const int64_t k = k_first + row;
    float2 *out = Z + (pair * nrows + ring_slot(ring0, row, nrows)) * kN;
    if (k < 0) {  // block before the start of the signal: all-zero spectrum
        for (int i = threadIdx.x; i < kN; i += kFftThreads) out[i] = make_float2(0.f, 0.f);
        return;
    }
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const float *xa = x + ca * ldx;
    const float *xb = x + (cb < C ? cb : ca) * ldx;
    const bool has_b = cb < C;
    const int64_t nbase = (k - 1) * kB + threadIdx.x;
    float2 v[16];
    if (nbase >= 0 && nbase + 256 * 15 < T) {  // whole block inside the signal (block-uniform up to 255 samples)
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = make_float2(__ldg(&xa[nbase + 256 * m]), __ldg(&xb[nbase + 256 * m]));
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int64_t n = nbase + 256 * m;
            const bool ok = n >= 0 && n < T;
            v[m] = make_float2(ok ? __ldg(&xa[n]) : 0.f, ok ? __ldg(&xb[n]) : 0.f);
        }
    }
    if (!has_b) {
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m].y = 0.f;
    }
    fft_fwd_tail(v, s, tab, out);
}

// ------------------------------------------------------------------------------------------
// Longer transforms, N = 4096 R (R = 2, 4): one radix-R decimation-in-frequency stage in front of
// R independent 4096-point transforms.  With a[n + 4096 j] the j-th quarter (half) of the block,
//     U_r = FFT_4096( W_N^{r n} * sum_j a[n + 4096 j] * w_R^{j r} ),  w_R = exp(-2 pi i / R),
// is the decimated spectrum X[R k + r]; the R sub-spectra are stored one after the other, each in
// the 4096-point kernel's own output order -- again only ever multiplied point-wise.  Forward: one
// CTA per (block, pair, r), the R x 16 inputs of a thread are combined in registers (the re-reads of
// x hit L2).  Inverse: one CTA runs the R sub-transforms in turn and accumulates only the VALID
// half of the block, a[n + 4096 j] = (1/N) sum_r conj(w_R^{j r} W_N^{r n}) u_r[n] for j >= R/2.
// Why: the frequency-domain delay line costs P = K / (2048 R) complex MACs per bin and the
// transforms ~log N per sample (see pick_r for what that measured on B200: not a win in this form).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 mul_mi_pow(float2 v, int e) {  // v * (-i)^e
    switch (e & 3) {
        case 1: return make_float2(v.y, -v.x);
        case 2: return make_float2(-v.x, -v.y);
        case 3: return make_float2(-v.y, v.x);
        default: return v;
    }
}
// twn[k] = exp(-2 pi i k / N), k < N
__global__ void fir_twiddle_n_kernel(float2 *twn, int n_fft) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_fft) {
        double sn, cs;
        sincospi(-2.0 * k / n_fft, &sn, &cs);
        twn[k] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
    }
}

template <int R>
__global__ void __launch_bounds__(kFftThreads) fir_taps_fft_r_kernel(const float *__restrict__ taps, int64_t K, float2 *__restrict__ H,
                                                                    const float4 *__restrict__ tab, const float2 *__restrict__ twn) {
    __shared__ float2 s[kPadN];
    constexpr int N = kN * R, B = kB * R;
    const int64_t p = blockIdx.x;
    const int r = blockIdx.y;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = threadIdx.x + 256 * m;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {  // the partition fills the first half of the block only
            const int i = n + kN * j;
            const int64_t t = p * B + i;
            const float2 a = make_float2(t < K ? __ldg(&taps[t]) : 0.f, 0.f);
            const float2 w = mul_mi_pow(a, (4 / R) * j * r);
            acc.x += w.x;
            acc.y += w.y;
        }
        v[m] = r > 0 ? cmul(acc, __ldg(&twn[r * n])) : acc;
    }
    fft_fwd_tail(v, s, tab, H + p * N + r * kN);
}

template <int R>
__global__ void __launch_bounds__(kFftThreads) fir_fwd_r_kernel(const float *__restrict__ x, int64_t C, int64_t T, int64_t ldx,
                                                               int64_t k_first, int64_t nrows, int64_t row0, int64_t ring0,
                                                               float2 *__restrict__ Z, const float4 *__restrict__ tab,
                                                               const float2 *__restrict__ twn) {
    __shared__ float2 s[kPadN];
    constexpr int N = kN * R, B = kB * R;
    const int64_t row = row0 + blockIdx.x;
    const int64_t pair = blockIdx.y;
    const int r = blockIdx.z;
    const int64_t k = k_first + row;
    float2 *out = Z + (pair * nrows + ring_slot(ring0, row, nrows)) * N + r * kN;
    if (k < 0) {  // block before the start of the signal: all-zero spectrum
        for (int i = threadIdx.x; i < kN; i += kFftThreads) out[i] = make_float2(0.f, 0.f);
        return;
    }
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const float *xa = x + ca * ldx;
    const float *xb = x + (cb < C ? cb : ca) * ldx;
    const bool has_b = cb < C;
    const int64_t nbase = (k - 1) * B + threadIdx.x;
    const bool inside = nbase >= 0 && nbase + N - 1 - threadIdx.x < T;  // block-uniform: the whole block is inside the signal
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = threadIdx.x + 256 * m;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t t = nbase + 256 * m + kN * j;
            float2 a;
            if (inside) {
                a = make_float2(__ldg(&xa[t]), __ldg(&xb[t]));
            } else {
                const bool ok = t >= 0 && t < T;
                a = make_float2(ok ? __ldg(&xa[t]) : 0.f, ok ? __ldg(&xb[t]) : 0.f);
            }
            if (!has_b) a.y = 0.f;
            const float2 w = mul_mi_pow(a, (4 / R) * j * r);
            acc.x += w.x;
            acc.y += w.y;
        }
        v[m] = r > 0 ? cmul(acc, __ldg(&twn[r * n])) : acc;
    }
    fft_fwd_tail(v, s, tab, out);
}

template <int R>
__global__ void __launch_bounds__(kFftThreads, R == 2 ? 2 : 1) fir_inv_r_kernel(const float2 *__restrict__ Y, float *__restrict__ y, int64_t C, int64_t T,
                                                               int64_t ldy, int64_t k_first, int64_t nout,
                                                               const float4 *__restrict__ tab, const float2 *__restrict__ twn) {
    __shared__ float2 s[kPadN];
    constexpr int N = kN * R, B = kB * R, HV = R / 2;  // HV valid quarters (halves) of kN samples each
    const int64_t jb = blockIdx.x;
    const int64_t pair = blockIdx.y;
    float2 acc[HV][16];
#pragma unroll
    for (int h = 0; h < HV; ++h)
#pragma unroll
        for (int m = 0; m < 16; ++m) acc[h][m] = make_float2(0.f, 0.f);
    float2 v[16];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r > 0) __syncthreads();  // everyone is done reading the previous sub-transform's tile
        fft_inv(v, s, tab, Y + (pair * nout + jb) * N + r * kN);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int n = threadIdx.x + 256 * m;
            const float2 t = r > 0 ? cmulc(v[m], __ldg(&twn[r * n])) : v[m];
#pragma unroll
            for (int h = 0; h < HV; ++h) {
                const float2 w = mul_mi_pow(t, -(4 / R) * (h + HV) * r);  // conj(w_R^{j r}) = (-i)^{-(4/R) j r}
                acc[h][m].x += w.x;
                acc[h][m].y += w.y;
            }
        }
    }
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const float scale = 1.0f / N;
    float *ya = y + ca * ldy;
    float *yb = y + cb * ldy;
    const bool has_b = cb < C;
#pragma unroll
    for (int h = 0; h < HV; ++h)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int64_t n = (k_first + jb) * B + h * kN + threadIdx.x + 256 * m;
            if (n < T) {
                ya[n] = acc[h][m].x * scale;
                if (has_b) yb[n] = acc[h][m].y * scale;
            }
        }
}

#if !TFX_MAC_BULK
// 16-byte cp.async that writes zeros instead when !valid (src-size 0: nothing is read).
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 16 : 0)
                 : "memory");
}
#endif

// Y[pair][j] = sum_p H[p] . Z[pair][j + (P-1) - p],  j in [0, nout)
//
// Persistent CTAs (one per SM, 512 threads) walk the (pair, 128-block, 64-bin) tiles; a tile with
// P > 32 partitions is a run of 32-partition work items that accumulate in registers.  The
// [159 x 64] spectra rows and [32 x 64] taps of item i+1 are fetched with cp.async into the other
// half of a double buffer while item i is multiplied, so the FMA pipe never waits on a staging phase
// (the first version loaded, synchronised and then computed with two CTAs per SM: 39 % FMA
// utilisation, `long_scoreboard` the dominant stall).
// Inner loop: each thread owns one bin and kMacPerThread = 16 consecutive blocks and slides a
// register window over the rows: 2 shared loads per 64 FFMA.
struct MacItem {
    int pair, j0, f0, pc;
};
__global__ void __launch_bounds__(512, 1) fir_mac_kernel(const float2 *__restrict__ Z, const float2 *__restrict__ H,
                                                        float2 *__restrict__ Y, int P, int nrows, int ring0, int nvalid,
                                                        int nout, int ntiles_j, int ntiles, int n_fft) {
    extern __shared__ float2 smc[];
    constexpr int kBufElems = (kMacRows + kMacPc) * kMacBins;  // Zs [159][64] then Hs [32][64]
    constexpr int R = kMacPerThread;
    const int fl = threadIdx.x & (kMacBins - 1);
    const int kg = threadIdx.x >> 6;  // 0..7: which R consecutive blocks
    const int nchunks = (P + kMacPc - 1) / kMacPc;

    auto decode = [&](int t, int pc) {
        MacItem it;
        it.pc = pc;
        it.f0 = (t % (n_fft / kMacBins)) * kMacBins;
        t /= (n_fft / kMacBins);
        it.j0 = (t % ntiles_j) * kMacBlocks;
        it.pair = t / ntiles_j;
        return it;
    };
    // partitions of chunk pc that exist, rounded up to the unroll step of the short path
    auto chunk_parts = [&](int pc) { return min(kMacPc, ((P - pc * kMacPc) + 7) & ~7); };
#if TFX_MAC_BULK
    // Staging by the bulk-copy engine: one 512-byte `cp.async.bulk` per spectra / taps row, issued by the lanes
    // of warp 0 and counted on the buffer's mbarrier -- ~12 instructions per item instead of ~300 per thread
    // (the LDGSTS version spent a quarter of its issue slots on staging addresses).  Rows outside the signal
    // or beyond the last partition are zero-filled with ordinary stores.
    uint64_t *bars = reinterpret_cast<uint64_t *>(smc + 2 * kBufElems);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_proxy_async_smem();
    }
    __syncthreads();
    auto issue = [&](const MacItem &it, float2 *buf, uint64_t *bar) {
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        const int npl = chunk_parts(it.pc);
        const float2 *Zp = Z + static_cast<int64_t>(it.pair) * nrows * n_fft + it.f0;
        const int row_lo = it.j0 + (P - 1) - (it.pc * kMacPc + kMacPc - 1);
        const int r_first = kMacPc - npl;  // local rows below this belong to partitions that are not run
        const int n_ok = max(min(row_lo + kMacRows, nvalid) - max(row_lo + r_first, 0), 0);
        const int h_ok = max(min(npl, P - it.pc * kMacPc), 0);
        if (lane == 0) mbar_arrive_expect_tx(bar, 512u * static_cast<uint32_t>(n_ok + h_ok));
        fence_proxy_async_smem();
        __syncwarp();
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r_first + lane; r < kMacRows; r += 32) {
            const int row = row_lo + r;
            float2 *dst = buf + r * kMacBins;
            if (row >= 0 && row < nvalid) {
                int slot = ring0 + row;
                slot = slot >= nrows ? slot - nrows : slot;
                bulk_load_1d(dst, Zp + static_cast<int64_t>(slot) * n_fft, 512u, bar);
            } else {
                for (int q = 0; q < 32; ++q) reinterpret_cast<float4 *>(dst)[q] = zero;
            }
        }
        float2 *hb = buf + kMacRows * kMacBins;
        for (int pl = lane; pl < npl; pl += 32) {
            const int p = it.pc * kMacPc + pl;
            float2 *dst = hb + pl * kMacBins;
            if (p < P) {
                bulk_load_1d(dst, H + static_cast<int64_t>(p) * n_fft + it.f0, 512u, bar);
            } else {
                for (int q = 0; q < 32; ++q) reinterpret_cast<float4 *>(dst)[q] = zero;
            }
        }
    };
    unsigned phase = 0;  // bit b: parity to wait for on buffer b
#else
    auto issue = [&](const MacItem &it, float2 *buf) {
        const int npl = chunk_parts(it.pc);
        const float2 *Zp = Z + static_cast<int64_t>(it.pair) * nrows * n_fft + it.f0;
        // rows needed: j + (P-1) - p for j in [j0, j0 + kMacBlocks), p in [32pc, 32pc+npl)
        const int row_lo = it.j0 + (P - 1) - (it.pc * kMacPc + kMacPc - 1);
        const int r_first = kMacPc - npl;  // local rows below this belong to partitions that are not run
        // Thread (w, q) copies 16-byte piece q of local rows r_first + w, + 16, ...: ring slot, source pointer and
        // shared-memory address advance by constants (the staging loop used to redo the ring arithmetic and a
        // 64-bit multiply per row -- integer work that competes with the FFMAs for the FMA pipe).
        const int q = threadIdx.x & 31;
        int r = (threadIdx.x >> 5) + r_first;
        int row = row_lo + r;
        int slot = ring0 + row;  // may be negative while row < 0: never dereferenced then
        if (slot >= nrows) slot -= nrows;
        const float2 *src = Zp + static_cast<int64_t>(slot) * n_fft + 2 * q;
        const int64_t step = static_cast<int64_t>(16) * n_fft, wrap = static_cast<int64_t>(nrows) * n_fft;
        float2 *dst = buf + r * kMacBins + 2 * q;
#pragma unroll 2
        for (; r < kMacRows; r += 16) {
            const bool ok = static_cast<unsigned>(row) < static_cast<unsigned>(nvalid);
            cp_async16_zfill(dst, ok ? src : Zp, ok);
            row += 16;
            slot += 16;
            src += step;
            if (slot >= nrows) {
                slot -= nrows;
                src -= wrap;
            }
            dst += 16 * kMacBins;
        }
        float2 *hb = buf + kMacRows * kMacBins;
        for (int i = threadIdx.x; i < npl * 32; i += 512) {
            const int pl = i >> 5;
            const int p = it.pc * kMacPc + pl;
            const bool ok = p < P;
            cp_async16_zfill(hb + pl * kMacBins + 2 * q, H + (ok ? static_cast<int64_t>(p) * n_fft : 0) + it.f0 + 2 * q, ok);
        }
        cp_async_commit();
    };
#endif
    float2 acc[R];
    // this CTA's work: tiles blockIdx.x, blockIdx.x + gridDim.x, ...; within a tile chunks 0..nchunks-1
    int tile = blockIdx.x;
    int pc = 0;
    if (tile >= ntiles) return;
    MacItem cur = decode(tile, pc);
    int b = 0;
#if TFX_MAC_BULK
    issue(cur, smc, &bars[0]);
    __syncthreads();  // zero-filled rows of the first item are visible to everyone
#else
    issue(cur, smc);
#endif
    while (tile < ntiles) {
        int ntile = tile;
        int npc = pc + 1;
        if (npc == nchunks) {
            npc = 0;
            ntile = tile + gridDim.x;
        }
        MacItem nxt = cur;
#if TFX_MAC_BULK
        if (ntile < ntiles) {
            nxt = decode(ntile, npc);
            issue(nxt, smc + (b ^ 1) * kBufElems, &bars[b ^ 1]);
        }
        mbar_wait(&bars[b], (phase >> b) & 1u);
        phase ^= 1u << b;
#else
        if (ntile < ntiles) {
            nxt = decode(ntile, npc);
            issue(nxt, smc + (b ^ 1) * kBufElems);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#endif
        if (cur.pc == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
        }
        {
            const float2 *Zs = smc + b * kBufElems;
            const float2 *Hs = Zs + kMacRows * kMacBins;
            // local row of (block jj, partition pl) = kg*R + jj + 31 - pl = m0 + jj - pl, m0 = kg*R + 31
            const float2 *zcol = Zs + fl;
            const int m0 = kg * R + (kMacPc - 1);
            const int npl = chunk_parts(cur.pc);
            float2 w[R];  // w[m & (R-1)] = Zs[m0 + m - pl] window, m = jj
#pragma unroll
            for (int m = 0; m < R; ++m) w[m] = zcol[(m0 + m) * kMacBins];
            if (npl == kMacPc) {
#pragma unroll
                for (int pl = 0; pl < kMacPc; ++pl) {
                    const float2 h = Hs[pl * kMacBins + fl];
#pragma unroll
                    for (int jj = 0; jj < R; ++jj) {
                        const float2 z = w[(jj - pl) & (R - 1)];
                        acc[jj].x = fmaf(h.x, z.x, acc[jj].x);
                        acc[jj].x = fmaf(-h.y, z.y, acc[jj].x);
                        acc[jj].y = fmaf(h.x, z.y, acc[jj].y);
                        acc[jj].y = fmaf(h.y, z.x, acc[jj].y);
                    }
                    if (pl + 1 < kMacPc) w[(R - 1 - pl) & (R - 1)] = zcol[(m0 - pl - 1) * kMacBins];
                }
            } else {
                // short chunk (P not a multiple of 32): 8 partitions per unrolled step keep the window
                // indices compile-time; P = 1 costs 1/4 of a full chunk instead of all of it
                for (int pb = 0; pb < npl; pb += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int pl = pb + u;
                        const float2 h = Hs[pl * kMacBins + fl];
#pragma unroll
                        for (int jj = 0; jj < R; ++jj) {
                            const float2 z = w[(jj - u) & (R - 1)];
                            acc[jj].x = fmaf(h.x, z.x, acc[jj].x);
                            acc[jj].x = fmaf(-h.y, z.y, acc[jj].x);
                            acc[jj].y = fmaf(h.x, z.y, acc[jj].y);
                            acc[jj].y = fmaf(h.y, z.x, acc[jj].y);
                        }
                        if (pl + 1 < npl) w[(R - 1 - u) & (R - 1)] = zcol[(m0 - pl - 1) * kMacBins];
                    }
                    // after 8 partitions the window has slid 8 rows: rotate the register names back
                    // the window has slid 8 rows but holds 16: reload it rather than rotate register names
                    if (pb + 8 < npl) {
#pragma unroll
                        for (int m = 0; m < R; ++m) w[m] = zcol[(m0 + m - (pb + 8)) * kMacBins];
                    }
                }
            }
        }
        if (cur.pc == nchunks - 1) {
            float2 *Yp = Y + static_cast<int64_t>(cur.pair) * nout * n_fft + cur.f0 + fl;
#pragma unroll
            for (int jj = 0; jj < R; ++jj) {
                const int j = cur.j0 + kg * R + jj;
                if (j < nout) Yp[static_cast<int64_t>(j) * n_fft] = acc[jj];
            }
        }
        __syncthreads();  // buffer b is free for the fetch issued in the next iteration (and the other buffer's zero rows are visible)
        cur = nxt;
        tile = ntile;
        pc = npc;
        b ^= 1;
    }
}

// y[kB : (k+1)B] of both channels of the pair = IFFT(Y[pair][j])[B : 2B] / N,  k = k_first + j
__global__ void __launch_bounds__(kFftThreads) fir_inv_kernel(const float2 *__restrict__ Y, float *__restrict__ y, int64_t C, int64_t T,
                                                             int64_t ldy, int64_t k_first, int64_t nout,
                                                             const float4 *__restrict__ tab) {
    __shared__ float2 s[kPadN];
    const int64_t j = blockIdx.x;
    const int64_t pair = blockIdx.y;
    float2 v[16];
    fft_inv(v, s, tab, Y + (pair * nout + j) * kN);
    // v[m] = time sample threadIdx.x + 256 m of the block; the valid half is m >= 8
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const int64_t nbase = (k_first + j) * kB + threadIdx.x;
    const float scale = 1.0f / kN;
    float *ya = y + ca * ldy;
    float *yb = y + cb * ldy;
    const bool has_b = cb < C;
#pragma unroll
    for (int m = 8; m < 16; ++m) {
        const int64_t n = nbase + 256 * (m - 8);
        if (n < T) {
            ya[n] = v[m].x * scale;
            if (has_b) yb[n] = v[m].y * scale;
        }
    }
}

struct OlsLayout {
    int R;  // transform size N = 4096 R, partition / hop B = 2048 R
    int64_t N, B;
    int64_t P, npairs, nblk, slab, nrows;  // slab = output blocks per slab, nrows = slab + P - 1
    size_t off_tw, off_twn, off_H, off_Z, off_Y, total;
};

// Transform size.  The delay line costs P = K / (2048 R) complex MACs per bin, so a longer transform
// halves / quarters the MAC work -- but measured on B200 for the 65 536-tap config (profiles/r1_fir.md)
// R = 2 and R = 4 LOSE (8.8 -> 10.2 -> 14.0 ms): every sub-transform CTA re-reads all R parts of the block
// (forward kernel 2.2x / 4x slower), the register-heavy inverse runs at 1-2 CTAs per SM, and the MAC, now
// with half the flops, drops onto its own HBM bound.  R = 1 therefore stays the default for every K;
// TFX_FIR_R = 2 | 4 keeps the longer transforms reachable (parity-tested) for further work.
int pick_r(int64_t K) {
    (void)K;
    if (const char *e = std::getenv("TFX_FIR_R")) {
        const int r = std::atoi(e);
        if (r == 1 || r == 2 || r == 4) return r;
    }
    return 1;
}

OlsLayout ols_layout(int64_t C, int64_t T, int64_t K) {
    OlsLayout L{};
    L.R = pick_r(K);
    L.N = static_cast<int64_t>(kN) * L.R;
    L.B = static_cast<int64_t>(kB) * L.R;
    L.P = (K + L.B - 1) / L.B;
    L.npairs = (C + 1) / 2;
    L.nblk = (T + L.B - 1) / L.B;
    // Slab: as many blocks as keep Z + Y near 1 GiB, at least 4P so the ring's P-1 history rows
    // stay a small fraction, and at least one MAC tile.
    const int64_t per_block = L.npairs * L.N * static_cast<int64_t>(sizeof(float2)) * 2;
    int64_t slab = (int64_t(1) << 30) / std::max<int64_t>(per_block, 1);
    slab = std::max<int64_t>(slab, 4 * L.P);
    slab = std::max<int64_t>(slab, kMacBlocks);
    slab = std::min<int64_t>(slab, L.nblk);
    slab = std::max<int64_t>(slab, 1);
    L.slab = slab;
    L.nrows = slab + L.P - 1;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    L.off_tw = 0;
    L.off_twn = align(kTwBytes);
    L.off_H = align(L.off_twn + static_cast<size_t>(L.N) * sizeof(float2));
    L.off_Z = align(L.off_H + static_cast<size_t>(L.P) * L.N * sizeof(float2));
    L.off_Y = align(L.off_Z + static_cast<size_t>(L.npairs) * L.nrows * L.N * sizeof(float2));
    L.total = align(L.off_Y + static_cast<size_t>(L.npairs) * L.slab * L.N * sizeof(float2));
    return L;
}

// The first overlap-save implementation (4096-point transforms, three kernels per slab, spectra through HBM)
// stays reachable with TFX_FIR_V1=1 for A/B measurements; the default is the persistent kernel of fir_ols16k.cu.
bool use_v1() {
    const char *e = std::getenv("TFX_FIR_V1");
    return e != nullptr && e[0] == '1';
}

int pick_algo(int algo, int64_t K) {
    if (algo == TFX_FIR_DIRECT || algo == TFX_FIR_OLS) return algo;
    return K <= kAutoDirectTaps ? TFX_FIR_DIRECT : TFX_FIR_OLS;
}

}  // namespace
}  // namespace tfx

extern "C" {

size_t tfx_fir_workspace_bytes(int64_t C, int64_t T, int64_t K, int algo) {
    if (C <= 0 || T <= 0 || K <= 0) return 0;
    if (tfx::pick_algo(algo, K) == TFX_FIR_DIRECT && K <= tfx::kDirectMaxTaps) return 0;
    if (!tfx::use_v1()) return tfx::fir_ols16k_workspace_bytes(K);
    return tfx::ols_layout(C, T, K).total;
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
    int use = pick_algo(algo, K);
    if (use == TFX_FIR_DIRECT && K > kDirectMaxTaps) use = TFX_FIR_OLS;

    if (use == TFX_FIR_DIRECT) {
        const int64_t tiles = (T + kDirTile - 1) / kDirTile;
        const size_t smem = sizeof(float) * (((K + 3) & ~3) + kDirTile + K - 1);
        TFX_REQUIRE(C * tiles < (int64_t(1) << 31), "fir: too many tiles for one launch");
        fir_direct_kernel<<<static_cast<unsigned>(C * tiles), 256, smem, stream>>>(x, y, C, T, ldx, ldy, taps, static_cast<int>(K),
                                                                                 tiles);
        TFX_CHECK_LAUNCH("fir_direct_kernel");
        return TFX_OK;
    }

    if (!use_v1()) return launch_fir_ols16k(x, y, C, T, ldx, ldy, taps, K, workspace, workspace_bytes, stream);

    const OlsLayout L = ols_layout(C, T, K);
    if (workspace == nullptr || workspace_bytes < L.total) {
        set_error("fir: workspace of %zu bytes needed, %zu given (query tfx_fir_workspace_bytes)", L.total, workspace_bytes);
        return TFX_EWORKSPACE;
    }
    TFX_REQUIRE(L.P <= 65535 && L.npairs <= 65535, "fir: too many partitions / channel pairs for one launch");
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float2 *tw2 = reinterpret_cast<float2 *>(ws + L.off_tw);
    const float4 *tw = reinterpret_cast<const float4 *>(tw2);
    float2 *twn = reinterpret_cast<float2 *>(ws + L.off_twn);
    float2 *H = reinterpret_cast<float2 *>(ws + L.off_H);
    float2 *Z = reinterpret_cast<float2 *>(ws + L.off_Z);
    float2 *Y = reinterpret_cast<float2 *>(ws + L.off_Y);
    const unsigned P = static_cast<unsigned>(L.P), npairs = static_cast<unsigned>(L.npairs);
    fir_twiddle_kernel<<<(kTwEntries + 255) / 256, 256, 0, stream>>>(tw2);
    TFX_CHECK_LAUNCH("fir_twiddle_kernel");
    if (L.R > 1) {
        fir_twiddle_n_kernel<<<static_cast<unsigned>((L.N + 255) / 256), 256, 0, stream>>>(twn, static_cast<int>(L.N));
        TFX_CHECK_LAUNCH("fir_twiddle_n_kernel");
    }
    if (L.R == 1)
        fir_taps_fft_kernel<<<P, kFftThreads, 0, stream>>>(taps, K, H, tw);
    else if (L.R == 2)
        fir_taps_fft_r_kernel<2><<<dim3(P, 2), kFftThreads, 0, stream>>>(taps, K, H, tw, twn);
    else
        fir_taps_fft_r_kernel<4><<<dim3(P, 4), kFftThreads, 0, stream>>>(taps, K, H, tw, twn);
    TFX_CHECK_LAUNCH("fir_taps_fft_kernel");
    const size_t mac_smem = 2 * sizeof(float2) * (kMacRows + kMacPc) * kMacBins + 16;  // double buffer + two mbarriers
    TFX_ENSURE_SMEM(fir_mac_kernel, static_cast<int>(mac_smem));
    for (int64_t k0 = 0; k0 < L.nblk; k0 += L.slab) {
        const int64_t nout = std::min<int64_t>(L.slab, L.nblk - k0);
        const int64_t nrows = L.nrows;  // row pitch of Z per pair, the same for every slab
        const int64_t k_first = k0 - (L.P - 1);
        // The P-1 history spectra of this slab are the last P-1 of the previous one and are still in the
        // ring: only the new blocks are transformed.  The first slab's history (k < 0) is written as zeros.
        const int64_t row0 = k0 > 0 ? L.P - 1 : 0;
        const int64_t ring0 = k0 % nrows;
        const unsigned nfwd = static_cast<unsigned>(nout + L.P - 1 - row0);
        if (L.R == 1)
            fir_fwd_kernel<<<dim3(nfwd, npairs), kFftThreads, 0, stream>>>(x, C, T, ldx, k_first, nrows, row0, ring0, Z, tw);
        else if (L.R == 2)
            fir_fwd_r_kernel<2><<<dim3(nfwd, npairs, 2), kFftThreads, 0, stream>>>(x, C, T, ldx, k_first, nrows, row0, ring0, Z, tw, twn);
        else
            fir_fwd_r_kernel<4><<<dim3(nfwd, npairs, 4), kFftThreads, 0, stream>>>(x, C, T, ldx, k_first, nrows, row0, ring0, Z, tw, twn);
        TFX_CHECK_LAUNCH("fir_fwd_kernel");
        const int64_t ntiles_j = (nout + kMacBlocks - 1) / kMacBlocks;
        const int64_t ntiles = L.npairs * ntiles_j * (L.N / kMacBins);
        TFX_REQUIRE(ntiles < (int64_t(1) << 31) && nrows < (int64_t(1) << 19), "fir: slab too large for the MAC kernel's 32-bit indexing");
        const unsigned mac_grid = static_cast<unsigned>(std::min<int64_t>(ntiles, sm_count()));
        fir_mac_kernel<<<mac_grid, 512, mac_smem, stream>>>(Z, H, Y, static_cast<int>(L.P), static_cast<int>(nrows),
                                                          static_cast<int>(ring0), static_cast<int>(nout + L.P - 1),
                                                          static_cast<int>(nout), static_cast<int>(ntiles_j),
                                                          static_cast<int>(ntiles), static_cast<int>(L.N));
        TFX_CHECK_LAUNCH("fir_mac_kernel");
        const unsigned ninv = static_cast<unsigned>(nout);
        if (L.R == 1)
            fir_inv_kernel<<<dim3(ninv, npairs), kFftThreads, 0, stream>>>(Y, y, C, T, ldy, k0, nout, tw);
        else if (L.R == 2)
            fir_inv_r_kernel<2><<<dim3(ninv, npairs), kFftThreads, 0, stream>>>(Y, y, C, T, ldy, k0, nout, tw, twn);
        else
            fir_inv_r_kernel<4><<<dim3(ninv, npairs), kFftThreads, 0, stream>>>(Y, y, C, T, ldy, k0, nout, tw, twn);
        TFX_CHECK_LAUNCH("fir_inv_kernel");
    }
    return TFX_OK;
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
