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

#include "common.cuh"

namespace tfx {
namespace {

constexpr int kDirectMaxTaps = 1024;  // smem-limited; AUTO switches to OLS far earlier
constexpr int kAutoDirectTaps = 96;
constexpr int kN = 4096;              // complex FFT size
constexpr int kB = 2048;              // partition / hop
constexpr int kFftThreads = 256;
constexpr int kMacBins = 64;          // bins per MAC CTA
constexpr int kMacBlocks = 64;        // output blocks per MAC CTA
constexpr int kMacPc = 32;            // partitions per shared-memory chunk
constexpr int kMacRows = kMacBlocks + kMacPc - 1;  // 95 spectra rows staged per chunk
constexpr size_t kTwBytes = sizeof(float2) * kN;

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

// forward, decimation in frequency: natural order in, base-4 digit-reversed order out.
// Pass with stride q (= 256, 16, 1) fuses the radix-4 stages of quarter 4q and q.
__device__ void fft_dif4(float2 *s, const float2 *__restrict__ tw) {
#pragma unroll 1
    for (int lq = 8; lq >= 0; lq -= 4) {
        const int q = 1 << lq;
        const int j = threadIdx.x & (q - 1);
        const int base = ((threadIdx.x >> lq) << (lq + 4)) + j;
        // twiddle loads first (global, L1/L2): their latency overlaps the shared-memory loads
        const bool twb = lq > 0;
        const int ta = j << (8 - lq);   // j * N / (16q)
        const int tb = j << (10 - lq);  // j * N / (4q)
        const float2 one = make_float2(1.f, 0.f);
        const float2 a1 = twb ? __ldg(&tw[ta]) : one, a2 = twb ? __ldg(&tw[2 * ta]) : one, a3 = twb ? __ldg(&tw[3 * ta]) : one;
        const float2 w1 = twb ? __ldg(&tw[tb]) : one, w2 = twb ? __ldg(&tw[2 * tb]) : one, w3 = twb ? __ldg(&tw[3 * tb]) : one;
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = s[pidx(base + m * q)];
        // stage A: quarter 4q, butterflies (m, m+4, m+8, m+12), position jj = j + m*q
#pragma unroll
        for (int m = 0; m < 4; ++m)
            bfly4_dif(v[m], v[m + 4], v[m + 8], v[m + 12], true, m ? cmul(a1, w16(m)) : a1, m ? cmul(a2, w16(2 * m)) : a2,
                      m ? cmul(a3, w16(3 * m)) : a3);
        // stage B: quarter q, butterflies (4a, 4a+1, 4a+2, 4a+3), position j
#pragma unroll
        for (int a = 0; a < 4; ++a) bfly4_dif(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3], twb, w1, w2, w3);
#pragma unroll
        for (int m = 0; m < 16; ++m) s[pidx(base + m * q)] = v[m];
        __syncthreads();
    }
}

// inverse (unscaled), decimation in time: digit-reversed order in, natural order out
__device__ void fft_dit4_inv(float2 *s, const float2 *__restrict__ tw) {
#pragma unroll 1
    for (int lq = 0; lq <= 8; lq += 4) {
        const int q = 1 << lq;
        const int j = threadIdx.x & (q - 1);
        const int base = ((threadIdx.x >> lq) << (lq + 4)) + j;
        const bool twb = lq > 0;
        const int ta = j << (8 - lq);
        const int tb = j << (10 - lq);
        const float2 one = make_float2(1.f, 0.f);
        const float2 a1 = twb ? __ldg(&tw[ta]) : one, a2 = twb ? __ldg(&tw[2 * ta]) : one, a3 = twb ? __ldg(&tw[3 * ta]) : one;
        const float2 w1 = twb ? __ldg(&tw[tb]) : one, w2 = twb ? __ldg(&tw[2 * tb]) : one, w3 = twb ? __ldg(&tw[3 * tb]) : one;
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = s[pidx(base + m * q)];
        // stage B first (quarter q), then stage A (quarter 4q): the mirror of the forward pass
#pragma unroll
        for (int a = 0; a < 4; ++a) bfly4_dit_inv(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3], twb, w1, w2, w3);
#pragma unroll
        for (int m = 0; m < 4; ++m)
            bfly4_dit_inv(v[m], v[m + 4], v[m + 8], v[m + 12], true, m ? cmul(a1, w16(m)) : a1, m ? cmul(a2, w16(2 * m)) : a2,
                          m ? cmul(a3, w16(3 * m)) : a3);
#pragma unroll
        for (int m = 0; m < 16; ++m) s[pidx(base + m * q)] = v[m];
        __syncthreads();
    }
}

__global__ void fir_twiddle_kernel(float2 *tw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < kN) {
        double sn, cs;
        sincospi(-2.0 * t / kN, &sn, &cs);
        tw[t] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
    }
}

// H[p] = DIF-FFT(taps[pB:(p+1)B] zero-padded to N), one CTA per partition
__global__ void __launch_bounds__(kFftThreads) fir_taps_fft_kernel(const float *__restrict__ taps, int64_t K, float2 *__restrict__ H,
                                                                  const float2 *__restrict__ tw) {
    __shared__ float2 s[kPadN];
    const int64_t p = blockIdx.x;
    for (int i = threadIdx.x; i < kN; i += kFftThreads) {
        const int64_t j = p * kB + i;
        s[pidx(i)] = make_float2((i < kB && j < K) ? taps[j] : 0.f, 0.f);
    }
    __syncthreads();
    fft_dif4(s, tw);
    float2 *out = H + p * kN;
    for (int i = threadIdx.x; i < kN; i += kFftThreads) out[i] = s[pidx(i)];
}

// Z[pair][row] = DIF-FFT(x_a[(k-1)B : (k+1)B] + i x_b[...]),  k = k_first + row (k < 0 -> zeros)
// Rows [row0, row0 + gridDim.x) of the slab are computed; rows below row0 hold the previous slab's
// last P-1 spectra (copied there by the host loop), nrows is the row pitch of Z per pair.
__global__ void __launch_bounds__(kFftThreads) fir_fwd_kernel(const float *__restrict__ x, int64_t C, int64_t T, int64_t ldx,
                                                             int64_t k_first, int64_t nrows, int64_t row0,
                                                             float2 *__restrict__ Z, const float2 *__restrict__ tw) {
    __shared__ float2 s[kPadN];
    const int64_t row = row0 + blockIdx.x;
    const int64_t pair = blockIdx.y;
    const int64_t k = k_first + row;
    float2 *out = Z + (pair * nrows + row) * kN;
    if (k < 0) {  // block before the start of the signal: all-zero spectrum
        for (int i = threadIdx.x; i < kN; i += kFftThreads) out[i] = make_float2(0.f, 0.f);
        return;
    }
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const float *xa = x + ca * ldx;
    const float *xb = x + cb * ldx;
    const int64_t nbase = (k - 1) * kB;
    {
        float2 r[kN / kFftThreads];
#pragma unroll
        for (int u = 0; u < kN / kFftThreads; ++u) {
            const int64_t n = nbase + threadIdx.x + u * kFftThreads;
            const bool ok = n >= 0 && n < T;
            r[u] = make_float2(ok ? __ldg(&xa[n]) : 0.f, (ok && cb < C) ? __ldg(&xb[n]) : 0.f);
        }
#pragma unroll
        for (int u = 0; u < kN / kFftThreads; ++u) s[pidx(threadIdx.x + u * kFftThreads)] = r[u];
    }
    __syncthreads();
    fft_dif4(s, tw);
#pragma unroll
    for (int u = 0; u < kN / kFftThreads; ++u) out[threadIdx.x + u * kFftThreads] = s[pidx(threadIdx.x + u * kFftThreads)];
}

// Y[pair][j] = sum_p H[p] . Z[pair][j + (P-1) - p],  j in [0, nout)
__global__ void __launch_bounds__(512) fir_mac_kernel(const float2 *__restrict__ Z, const float2 *__restrict__ H,
                                                     float2 *__restrict__ Y, int P, int64_t nrows, int64_t nout) {
    extern __shared__ float2 smc[];
    float2 *Zs = smc;                        // [kMacRows][kMacBins]
    float2 *Hs = smc + kMacRows * kMacBins;  // [kMacPc][kMacBins]
    const int fl = threadIdx.x & (kMacBins - 1);
    const int kg = threadIdx.x >> 6;  // 0..7: which 8 consecutive blocks
    const int64_t f0 = static_cast<int64_t>(blockIdx.x) * kMacBins;
    const int64_t j0 = static_cast<int64_t>(blockIdx.y) * kMacBlocks;
    const int64_t pair = blockIdx.z;
    const float2 *Zp = Z + pair * nrows * kN;
    float2 acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = make_float2(0.f, 0.f);
    const int nchunks = (P + kMacPc - 1) / kMacPc;
    for (int pc = 0; pc < nchunks; ++pc) {
        // rows needed: j + (P-1) - p for j in [j0, j0+64), p in [32pc, 32pc+32)
        const int64_t row_lo = j0 + (P - 1) - (pc * kMacPc + kMacPc - 1);
        __syncthreads();
        for (int i = threadIdx.x; i < kMacRows * kMacBins; i += 512) {
            const int r = i >> 6, f = i & (kMacBins - 1);
            const int64_t row = row_lo + r;
            Zs[i] = (row >= 0 && row < nrows) ? Zp[row * kN + f0 + f] : make_float2(0.f, 0.f);
        }
        for (int i = threadIdx.x; i < kMacPc * kMacBins; i += 512) {
            const int pl = i >> 6, f = i & (kMacBins - 1);
            const int p = pc * kMacPc + pl;
            Hs[i] = p < P ? H[static_cast<int64_t>(p) * kN + f0 + f] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        // local row of (block jj, partition pl) = kg*8 + jj + 31 - pl = m0 + jj - pl, m0 = kg*8 + 31
        const float2 *zcol = Zs + fl;
        const int m0 = kg * 8 + (kMacPc - 1);
        float2 w[8];  // w[m & 7] = Zs[m0 + m - pl] window, m = jj
#pragma unroll
        for (int m = 0; m < 8; ++m) w[m] = zcol[(m0 + m) * kMacBins];
#pragma unroll
        for (int pl = 0; pl < kMacPc; ++pl) {
            const float2 h = Hs[pl * kMacBins + fl];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float2 z = w[(jj - pl) & 7];
                acc[jj].x = fmaf(h.x, z.x, acc[jj].x);
                acc[jj].x = fmaf(-h.y, z.y, acc[jj].x);
                acc[jj].y = fmaf(h.x, z.y, acc[jj].y);
                acc[jj].y = fmaf(h.y, z.x, acc[jj].y);
            }
            if (pl + 1 < kMacPc) w[(7 - pl) & 7] = zcol[(m0 - pl - 1) * kMacBins];
        }
    }
    float2 *Yp = Y + pair * nout * kN;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const int64_t j = j0 + kg * 8 + jj;
        if (j < nout) Yp[j * kN + f0 + fl] = acc[jj];
    }
}

// y[kB : (k+1)B] of both channels of the pair = IFFT(Y[pair][j])[B : 2B] / N,  k = k_first + j
__global__ void __launch_bounds__(kFftThreads) fir_inv_kernel(const float2 *__restrict__ Y, float *__restrict__ y, int64_t C, int64_t T,
                                                             int64_t ldy, int64_t k_first, int64_t nout,
                                                             const float2 *__restrict__ tw) {
    __shared__ float2 s[kPadN];
    const int64_t j = blockIdx.x;
    const int64_t pair = blockIdx.y;
    const float2 *in = Y + (pair * nout + j) * kN;
    {   // all 16 loads of a thread in flight before the first shared-memory store
        float2 r[kN / kFftThreads];
#pragma unroll
        for (int u = 0; u < kN / kFftThreads; ++u) r[u] = __ldcs(&in[threadIdx.x + u * kFftThreads]);
#pragma unroll
        for (int u = 0; u < kN / kFftThreads; ++u) s[pidx(threadIdx.x + u * kFftThreads)] = r[u];
    }
    __syncthreads();
    fft_dit4_inv(s, tw);
    const int64_t ca = 2 * pair, cb = 2 * pair + 1;
    const int64_t nbase = (k_first + j) * kB;
    const float scale = 1.0f / kN;
    for (int i = threadIdx.x; i < kB; i += kFftThreads) {
        const int64_t n = nbase + i;
        if (n < T) {
            const float2 v = s[pidx(kB + i)];
            y[ca * ldy + n] = v.x * scale;
            if (cb < C) y[cb * ldy + n] = v.y * scale;
        }
    }
}

struct OlsLayout {
    int64_t P, npairs, nblk, slab, nrows;  // slab = output blocks per slab, nrows = slab + P - 1
    size_t off_tw, off_H, off_Z, off_Y, total;
};

OlsLayout ols_layout(int64_t C, int64_t T, int64_t K) {
    OlsLayout L{};
    L.P = (K + kB - 1) / kB;
    L.npairs = (C + 1) / 2;
    L.nblk = (T + kB - 1) / kB;
    // Slab: as many blocks as keep Z + Y near 1 GiB, at least 4P so the P-1 recomputed
    // history blocks stay a small fraction.
    const int64_t per_block = L.npairs * static_cast<int64_t>(kN) * static_cast<int64_t>(sizeof(float2)) * 2;
    int64_t slab = (int64_t(1) << 30) / std::max<int64_t>(per_block, 1);
    slab = std::max<int64_t>(slab, 4 * L.P);
    slab = std::max<int64_t>(slab, kMacBlocks);
    slab = std::min<int64_t>(slab, L.nblk);
    slab = std::max<int64_t>(slab, 1);
    L.slab = slab;
    L.nrows = slab + L.P - 1;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    L.off_tw = 0;
    L.off_H = align(kTwBytes);
    L.off_Z = align(L.off_H + static_cast<size_t>(L.P) * kN * sizeof(float2));
    L.off_Y = align(L.off_Z + static_cast<size_t>(L.npairs) * L.nrows * kN * sizeof(float2));
    L.total = align(L.off_Y + static_cast<size_t>(L.npairs) * L.slab * kN * sizeof(float2));
    return L;
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

    const OlsLayout L = ols_layout(C, T, K);
    if (workspace == nullptr || workspace_bytes < L.total) {
        set_error("fir: workspace of %zu bytes needed, %zu given (query tfx_fir_workspace_bytes)", L.total, workspace_bytes);
        return TFX_EWORKSPACE;
    }
    TFX_REQUIRE(L.P <= 65535 && L.npairs <= 65535, "fir: too many partitions / channel pairs for one launch");
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float2 *tw = reinterpret_cast<float2 *>(ws + L.off_tw);
    float2 *H = reinterpret_cast<float2 *>(ws + L.off_H);
    float2 *Z = reinterpret_cast<float2 *>(ws + L.off_Z);
    float2 *Y = reinterpret_cast<float2 *>(ws + L.off_Y);
    fir_twiddle_kernel<<<kN / 256, 256, 0, stream>>>(tw);
    TFX_CHECK_LAUNCH("fir_twiddle_kernel");
    fir_taps_fft_kernel<<<static_cast<unsigned>(L.P), kFftThreads, 0, stream>>>(taps, K, H, tw);
    TFX_CHECK_LAUNCH("fir_taps_fft_kernel");
    const size_t mac_smem = sizeof(float2) * (kMacRows + kMacPc) * kMacBins;
    TFX_ENSURE_SMEM(fir_mac_kernel, static_cast<int>(mac_smem));
    const size_t row_bytes = sizeof(float2) * kN;
    for (int64_t k0 = 0; k0 < L.nblk; k0 += L.slab) {
        const int64_t nout = std::min<int64_t>(L.slab, L.nblk - k0);
        const int64_t nrows = L.nrows;  // row pitch of Z per pair, the same for every slab
        const int64_t k_first = k0 - (L.P - 1);
        // The P-1 history spectra of this slab are the last P-1 rows of the previous one: move them
        // to the front (device-to-device, ~1 % of a slab's traffic) instead of transforming the
        // same input blocks again; the first slab's history (k < 0) is written as zeros by the kernel.
        int64_t row0 = 0;
        if (k0 > 0 && L.P > 1) {
            row0 = L.P - 1;
            TFX_CUDA_TRY(cudaMemcpy2DAsync(Z, nrows * row_bytes, Z + L.slab * kN, nrows * row_bytes, row0 * row_bytes,
                                           static_cast<size_t>(L.npairs), cudaMemcpyDeviceToDevice, stream));
        }
        fir_fwd_kernel<<<dim3(static_cast<unsigned>(nout + L.P - 1 - row0), static_cast<unsigned>(L.npairs)), kFftThreads, 0,
                         stream>>>(x, C, T, ldx, k_first, nrows, row0, Z, tw);
        TFX_CHECK_LAUNCH("fir_fwd_kernel");
        fir_mac_kernel<<<dim3(kN / kMacBins, static_cast<unsigned>((nout + kMacBlocks - 1) / kMacBlocks),
                              static_cast<unsigned>(L.npairs)),
                         512, mac_smem, stream>>>(Z, H, Y, static_cast<int>(L.P), nrows, nout);
        TFX_CHECK_LAUNCH("fir_mac_kernel");
        fir_inv_kernel<<<dim3(static_cast<unsigned>(nout), static_cast<unsigned>(L.npairs)), kFftThreads, 0, stream>>>(
            Y, y, C, T, ldy, k0, nout, tw);
        TFX_CHECK_LAUNCH("fir_inv_kernel");
    }
    return TFX_OK;
}

}  // extern "C"
