// bank_stack.cu -- STACK filterbank (LogFilterBank, filter/filterbank.py:183-185: n_bands native
// calls + torch.stack) with lanes = channels.
//
// A CTA of W warps owns 32 CONSECUTIVE CHANNELS over one time segment.  The [32 ch x 64 samples]
// input tile is fetched ONCE per CTA (cp.async.cg, 128-byte XOR swizzle, two stages) and every
// warp filters it with "its" bands (band b belongs to warp b mod W): lane r runs the DF2T
// recurrence of band b over row r into the warp's private output tile, which then leaves the
// SM as 32 coalesced 256-byte rows of y[b, c0 .. c0+31, n .. n+63] (streaming 16-byte stores).
// So x costs 4/N bytes per lane-sample and y 4 bytes: the algorithmic 4*(1+1/N).
//
// Why not band-per-lane (bank_stream_kernel, filterbank.cu, still used for few channels and
// float64 I/O): with lanes = channels the band loop is warp-uniform, so
//   * precision is decided PER BAND (bit b of f64_mask): a 20 Hz band that needs the float64
//     recurrence no longer drags the 10 kHz bands onto the FP64 pipe;
//   * the warm-up launch (segment start states) runs each band only over ITS OWN decay length
//     (warm_b): in a log-spaced bank the sum of the decay lengths is ~5x the longest one, not Nx;
//   * band states live in shared memory between tiles ([slot][section][lane] double2, 2 LDS + 2 STS
//     per band per 64 samples) and coefficients come from the constant bank, so registers do not
//     limit the number of bands and all control flow and addressing is warp-uniform.
// SUM mode (`f1 + f2 + ...` with 9 .. 32 children, filter/__base.py:1019-1026): the same CTA, but a warp ADDS its
// bands into its tile (slot order), the warps' partial tiles are added in warp order after a barrier and the
// CTA stores one tile of y[C, T] per input tile -- 8 B per channel-sample whatever N is.
// DF1 state contract ([N, Kb, C, 2] float64, in place), time segmentation and warm-up launch are
// those of the other kernels (sos_plan.cpp); the last two samples of a channel run through a
// scalar epilogue that records each section's DF1 history.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "bank_tile.h"
#include "common.cuh"
#include "stream_common.cuh"

namespace tfx {
namespace {

#ifndef TFX_BS_WARPS
#define TFX_BS_WARPS 8
#endif
#ifndef TFX_BS_CH
#define TFX_BS_CH 64
#endif
#ifndef TFX_BS_STORE
#define TFX_BS_STORE 0
#endif
constexpr int kMaxWarps = TFX_BS_WARPS;
constexpr int kStagesX = 2;
constexpr int kCH = TFX_BS_CH;  // samples per chunk (float32 I/O): rows of 256 or 512 bytes
constexpr int kNV = kCH / 4;    // 16-byte pieces per row
constexpr int kTileBytes = 32 * kCH * 4;
constexpr int kMaxCtas = 8;
static_assert(kCH == 64 || kCH == 128, "tile rows are 256 or 512 bytes");
static_assert(32 % kMaxWarps == 0, "warps per CTA must divide 32");

template <int KB>
struct StackCoef {
    double b0[32][KB], b1[32][KB], b2[32][KB], a1[32][KB], a2[32][KB];
};

struct StackGeom {
    const float *x;
    float *y;
    int64_t ldx, ldy, ldb, C, T;
    int64_t S, Lseg, warm;  // warm > 0: this launch is the warm-up pass (warm = longest band warm-up)
    double *ws;             // [(b * KB + k) * 2 + h][C * S] segment start states (DF2T, double)
    int64_t ws_stride;
    double *state_x;  // [N, KB, C, 2]
    double *state_y;
    int n_bands;        // bands in this launch (<= 32)
    int W;              // warps per CTA
    int bpw;            // band slots per warp = ceil(n_bands / W)
    uint32_t f64_mask;  // bit b: band b runs the float64 recurrence
    int vec_ok;
    int nsplit;       // STACK: the bands are split over nsplit CTAs per (channel group, segment) -- more CTAs when C*T is small
    int bps;          // bands per split
    int sum;          // 1: SUM mode -- the bands are added (warp partials in slot order, then warps in order) into y[C, T]
    int band_id[32];  // global band index of local band b (y plane and state block)
    int warm_b[32];   // warm-up samples band b needs (multiple of 64, <= warm)
};

__host__ __device__ constexpr int warp_bytes(int bpw, int KB) { return kTileBytes + bpw * KB * 512; }
// input ring (float32) + one float64 copy of the current input tile + per-warp output tile and band states
__host__ __device__ constexpr int cta_bytes(int W, int bpw, int KB) { return (kStagesX + 2) * kTileBytes + W * warp_bytes(bpw, KB); }

// byte offset of the 16-byte column v (0..15) of row r inside a swizzled tile (as sos_tile.cuh)
__device__ __forceinline__ int col_offset(int r, int v) { return (v >> 3) * 4096 + r * 128 + (((v & 7) ^ (r & 7)) << 4); }
__device__ __forceinline__ void store_row16(float *gptr, const float4 &v) {
#if TFX_BS_STORE == 0
    st_stream16(gptr, v);
#else
    *reinterpret_cast<float4 *>(gptr) = v;
#endif
}
__device__ __forceinline__ int elem_offset(int r, int e) { return col_offset(r, e >> 2) + (e & 3) * 4; }

// One band over one chunk of my row: state in and out of shared memory.
// OUT: 0 = no output (warm-up), 1 = write the band's tile, 2 = add onto the tile (SUM mode, later bands of a warp)
__device__ __forceinline__ void emit4(unsigned char *p, const float4 &a, int out_mode) {
    if (out_mode == 1) {
        *reinterpret_cast<float4 *>(p) = a;
    } else if (out_mode == 2) {
        float4 o = *reinterpret_cast<const float4 *>(p);
        o.x += a.x;
        o.y += a.y;
        o.z += a.z;
        o.w += a.w;
        *reinterpret_cast<float4 *>(p) = o;
    }
}
__device__ __forceinline__ void emit1(unsigned char *p, float y, int out_mode) {
    if (out_mode == 1)
        *reinterpret_cast<float *>(p) = y;
    else if (out_mode == 2)
        *reinterpret_cast<float *>(p) += y;
}

// lo[j] = byte offset of 16-byte column j (j < 8) of THIS lane's row inside a swizzled tile; column v lives at
// lo[v & 7] + (v >> 3) * 4096, so the fully unrolled loops below address the tiles with immediates only.
template <typename CT, int KB, int OUT>
__device__ __forceinline__ void band_chunk(const StackCoef<KB> &cd, int b, const unsigned char *xin, unsigned char *out, double2 *st,
                                           int lane, int cnt, const int (&lo)[8]) {
    CT b0[KB], b1[KB], b2[KB], na1[KB], na2[KB], s1[KB], s2[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        b0[k] = static_cast<CT>(cd.b0[b][k]);
        b1[k] = static_cast<CT>(cd.b1[b][k]);
        b2[k] = static_cast<CT>(cd.b2[b][k]);
        na1[k] = static_cast<CT>(-cd.a1[b][k]);
        na2[k] = static_cast<CT>(-cd.a2[b][k]);
        const double2 s = st[k * 32 + lane];
        s1[k] = static_cast<CT>(s.x);
        s2[k] = static_cast<CT>(s.y);
    }
    auto step = [&](float x) -> float {
        CT v = static_cast<CT>(x);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const CT y = fma_rn(b0[k], v, s1[k]);
            s1[k] = fma_rn(na1[k], y, fma_rn(b1[k], v, s2[k]));
            s2[k] = fma_rn(na2[k], y, b2[k] * v);
            v = y;
        }
        return static_cast<float>(v);
    };
    if (cnt == kCH) {
        float4 a = *reinterpret_cast<const float4 *>(xin + lo[0]);
#pragma unroll
        for (int v = 0; v < kNV; ++v) {
            // next vector first: the compiler cannot hoist a load over the store into the output tile
            constexpr int kMask = kNV - 1;
            const int vn = (v + 1) & kMask;
            const float4 nxt = *reinterpret_cast<const float4 *>(xin + lo[vn & 7] + (vn >> 3) * 4096);
            a.x = step(a.x);
            a.y = step(a.y);
            a.z = step(a.z);
            a.w = step(a.w);
            emit4(out + lo[v & 7] + (v >> 3) * 4096, a, OUT);
            a = nxt;
        }
    } else {
        for (int e = 0; e < cnt; ++e) {
            const float y = step(*reinterpret_cast<const float *>(xin + elem_offset(lane, e)));
            emit1(out + elem_offset(lane, e), y, OUT);
        }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) st[k * 32 + lane] = make_double2(static_cast<double>(s1[k]), static_cast<double>(s2[k]));
}

// Float64 band: reads the CTA's float64 copy of the input tile (converted ONCE per tile, not once
// per band -- F2F runs at 16 lanes/clk/SM, so a per-band conversion costs the FP64 pipe more than
// the band's five DFMAs), rounds only the output.
template <int KB, int OUT>
__device__ __forceinline__ void band_chunk64(const StackCoef<KB> &cd, int b, const unsigned char *x64, unsigned char *out, double2 *st,
                                             int lane, int cnt, const int (&lo)[8]) {
    double b0[KB], b1[KB], b2[KB], na1[KB], na2[KB], s1[KB], s2[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        b0[k] = cd.b0[b][k];
        b1[k] = cd.b1[b][k];
        b2[k] = cd.b2[b][k];
        na1[k] = -cd.a1[b][k];
        na2[k] = -cd.a2[b][k];
        const double2 s = st[k * 32 + lane];
        s1[k] = s.x;
        s2[k] = s.y;
    }
    auto step = [&](double v) -> float {
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const double y = __fma_rn(b0[k], v, s1[k]);
            s1[k] = __fma_rn(na1[k], y, __fma_rn(b1[k], v, s2[k]));
            s2[k] = __fma_rn(na2[k], y, b2[k] * v);
            v = y;
        }
        return static_cast<float>(v);
    };
    if (cnt == kCH) {
        double2 a = *reinterpret_cast<const double2 *>(x64 + lo[0]);
#pragma unroll
        for (int v = 0; v < kNV; ++v) {
            constexpr int kMask = 2 * kNV - 1;
            const int c1 = 2 * v + 1, c2 = (2 * v + 2) & kMask;
            const double2 c = *reinterpret_cast<const double2 *>(x64 + lo[c1 & 7] + (c1 >> 3) * 4096);
            const double2 nxt = *reinterpret_cast<const double2 *>(x64 + lo[c2 & 7] + (c2 >> 3) * 4096);
            float4 o;
            o.x = step(a.x);
            o.y = step(a.y);
            o.z = step(c.x);
            o.w = step(c.y);
            emit4(out + lo[v & 7] + (v >> 3) * 4096, o, OUT);
            a = nxt;
        }
    } else {
        for (int e = 0; e < cnt; ++e) {
            const float y = step(*reinterpret_cast<const double *>(x64 + col_offset(lane, e >> 1) + (e & 1) * 8));
            emit1(out + elem_offset(lane, e), y, OUT);
        }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) st[k * 32 + lane] = make_double2(s1[k], s2[k]);
}

// TWO bands of the same precision over one chunk of my row (warm-up and SUM mode only: no per-band output tile
// is needed there).  The two recurrences are independent dependency chains, which doubles the instruction-level
// parallelism of a lane that otherwise waits ~8 cycles per sample on one chain.  Output: OUT = 0 none, 1 write
// y_a + y_b, 2 add y_a then y_b onto the tile (the warp's slot order is preserved).
template <typename CT, int KB, int OUT>
__device__ __forceinline__ void band_pair(const StackCoef<KB> &cd, int ba, int bb, const unsigned char *xsrc, unsigned char *out,
                                          double2 *sta, double2 *stb, int lane, int cnt, const int (&lo)[8]) {
    CT b0[2][KB], b1[2][KB], b2[2][KB], na1[2][KB], na2[2][KB], s1[2][KB], s2[2][KB];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int b = q ? bb : ba;
        const double2 *st = q ? stb : sta;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            b0[q][k] = static_cast<CT>(cd.b0[b][k]);
            b1[q][k] = static_cast<CT>(cd.b1[b][k]);
            b2[q][k] = static_cast<CT>(cd.b2[b][k]);
            na1[q][k] = static_cast<CT>(-cd.a1[b][k]);
            na2[q][k] = static_cast<CT>(-cd.a2[b][k]);
            const double2 s = st[k * 32 + lane];
            s1[q][k] = static_cast<CT>(s.x);
            s2[q][k] = static_cast<CT>(s.y);
        }
    }
    auto step2 = [&](CT x, float &ya, float &yb) {
        CT v[2] = {x, x};
#pragma unroll
        for (int k = 0; k < KB; ++k)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const CT y = fma_rn(b0[q][k], v[q], s1[q][k]);
                s1[q][k] = fma_rn(na1[q][k], y, fma_rn(b1[q][k], v[q], s2[q][k]));
                s2[q][k] = fma_rn(na2[q][k], y, b2[q][k] * v[q]);
                v[q] = y;
            }
        ya = static_cast<float>(v[0]);
        yb = static_cast<float>(v[1]);
    };
    auto emit = [&](unsigned char *p, const float4 &ya, const float4 &yb) {
        if (OUT == 1) {
            *reinterpret_cast<float4 *>(p) = make_float4(ya.x + yb.x, ya.y + yb.y, ya.z + yb.z, ya.w + yb.w);
        } else if (OUT == 2) {
            float4 o = *reinterpret_cast<const float4 *>(p);
            o.x = (o.x + ya.x) + yb.x;
            o.y = (o.y + ya.y) + yb.y;
            o.z = (o.z + ya.z) + yb.z;
            o.w = (o.w + ya.w) + yb.w;
            *reinterpret_cast<float4 *>(p) = o;
        }
    };
    if (cnt == kCH) {
#pragma unroll
        for (int v = 0; v < kNV; ++v) {
            CT x0, x1, x2, x3;
            if constexpr (sizeof(CT) == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(xsrc + lo[v & 7] + (v >> 3) * 4096);
                x0 = a.x;
                x1 = a.y;
                x2 = a.z;
                x3 = a.w;
            } else {
                constexpr int dummy = 0;
                (void)dummy;
                const int c0 = 2 * v, c1 = 2 * v + 1;
                const double2 p0 = *reinterpret_cast<const double2 *>(xsrc + lo[c0 & 7] + (c0 >> 3) * 4096);
                const double2 p1 = *reinterpret_cast<const double2 *>(xsrc + lo[c1 & 7] + (c1 >> 3) * 4096);
                x0 = p0.x;
                x1 = p0.y;
                x2 = p1.x;
                x3 = p1.y;
            }
            float4 ya, yb;
            step2(x0, ya.x, yb.x);
            step2(x1, ya.y, yb.y);
            step2(x2, ya.z, yb.z);
            step2(x3, ya.w, yb.w);
            emit(out + lo[v & 7] + (v >> 3) * 4096, ya, yb);
        }
    } else {
        for (int e = 0; e < cnt; ++e) {
            CT x;
            if constexpr (sizeof(CT) == 4)
                x = *reinterpret_cast<const float *>(xsrc + elem_offset(lane, e));
            else
                x = *reinterpret_cast<const double *>(xsrc + col_offset(lane, e >> 1) + (e & 1) * 8);
            float ya, yb;
            step2(x, ya, yb);
            if (OUT == 1)
                *reinterpret_cast<float *>(out + elem_offset(lane, e)) = ya + yb;
            else if (OUT == 2) {
                float *p = reinterpret_cast<float *>(out + elem_offset(lane, e));
                *p = (*p + ya) + yb;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        double2 *st = q ? stb : sta;
#pragma unroll
        for (int k = 0; k < KB; ++k) st[k * 32 + lane] = make_double2(static_cast<double>(s1[q][k]), static_cast<double>(s2[q][k]));
    }
}

// 2 * NP float32 bands over one chunk of my row on Blackwell's PACKED fp32 pipe (warm-up and SUM mode): the bands of slots
// (2q, 2q + 1) share one FFMA2 / FMUL2 per recurrence term -- coefficients and state packed as (band 2q, band 2q + 1), the
// sample broadcast into both halves.  The scalar pair kernel above is issue-bound (ncu, 32 bands x 256 ch: issue slots 67 %
// busy with the FMA pipe 50 % active, 27 % of the instructions are not math); packed math halves the issue slots of the math
// and leaves the FMA pipe as the bound.  The bands' outputs are added in slot order, as everywhere in SUM mode.
template <int KB, int NP, int OUT>
__device__ __forceinline__ void band_multi_p(const StackCoef<KB> &cd, const int (&bands)[2 * NP], const unsigned char *xsrc, unsigned char *out,
                                             double2 *const (&sts)[2 * NP], int lane, int cnt, const int (&lo)[8]) {
    float2 b0[NP][KB], b1[NP][KB], b2[NP][KB], na1[NP][KB], na2[NP][KB], s1[NP][KB], s2[NP][KB];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int ba = bands[2 * q], bb = bands[2 * q + 1];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            b0[q][k] = make_float2(static_cast<float>(cd.b0[ba][k]), static_cast<float>(cd.b0[bb][k]));
            b1[q][k] = make_float2(static_cast<float>(cd.b1[ba][k]), static_cast<float>(cd.b1[bb][k]));
            b2[q][k] = make_float2(static_cast<float>(cd.b2[ba][k]), static_cast<float>(cd.b2[bb][k]));
            na1[q][k] = make_float2(static_cast<float>(-cd.a1[ba][k]), static_cast<float>(-cd.a1[bb][k]));
            na2[q][k] = make_float2(static_cast<float>(-cd.a2[ba][k]), static_cast<float>(-cd.a2[bb][k]));
            const double2 sa = sts[2 * q][k * 32 + lane], sb = sts[2 * q + 1][k * 32 + lane];
            s1[q][k] = make_float2(static_cast<float>(sa.x), static_cast<float>(sb.x));
            s2[q][k] = make_float2(static_cast<float>(sa.y), static_cast<float>(sb.y));
        }
    }
    auto stepn = [&](float x) -> float {  // returns the slot-ordered sum of the 2 NP band outputs
        float2 v[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) v[q] = make_float2(x, x);
#pragma unroll
        for (int k = 0; k < KB; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const float2 y = __ffma2_rn(b0[q][k], v[q], s1[q][k]);
                s1[q][k] = __ffma2_rn(na1[q][k], y, __ffma2_rn(b1[q][k], v[q], s2[q][k]));
                s2[q][k] = __ffma2_rn(na2[q][k], y, __fmul2_rn(b2[q][k], v[q]));
                v[q] = y;
            }
        float t = v[0].x + v[0].y;
#pragma unroll
        for (int q = 1; q < NP; ++q) t = (t + v[q].x) + v[q].y;
        return t;
    };
    if (cnt == kCH) {
#pragma unroll
        for (int v = 0; v < kNV; ++v) {
            const float4 a = *reinterpret_cast<const float4 *>(xsrc + lo[v & 7] + (v >> 3) * 4096);
            float4 y;
            y.x = stepn(a.x);
            y.y = stepn(a.y);
            y.z = stepn(a.z);
            y.w = stepn(a.w);
            if (OUT == 1) {
                *reinterpret_cast<float4 *>(out + lo[v & 7] + (v >> 3) * 4096) = y;
            } else if (OUT == 2) {
                float4 o = *reinterpret_cast<const float4 *>(out + lo[v & 7] + (v >> 3) * 4096);
                o.x += y.x;
                o.y += y.y;
                o.z += y.z;
                o.w += y.w;
                *reinterpret_cast<float4 *>(out + lo[v & 7] + (v >> 3) * 4096) = o;
            }
        }
    } else {
        for (int e = 0; e < cnt; ++e) {
            const float y = stepn(*reinterpret_cast<const float *>(xsrc + elem_offset(lane, e)));
            if (OUT == 1)
                *reinterpret_cast<float *>(out + elem_offset(lane, e)) = y;
            else if (OUT == 2)
                *reinterpret_cast<float *>(out + elem_offset(lane, e)) += y;
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            sts[2 * q][k * 32 + lane] = make_double2(static_cast<double>(s1[q][k].x), static_cast<double>(s2[q][k].x));
            sts[2 * q + 1][k * 32 + lane] = make_double2(static_cast<double>(s1[q][k].y), static_cast<double>(s2[q][k].y));
        }
}

// The last `tail` (<= 2) samples of my channel for one band, straight from / to global memory,
// recording every section's input / output: that IS the DF1 state handed back.
template <typename CT, int KB>
__device__ __forceinline__ void band_tail(const StackCoef<KB> &cd, const StackGeom &g, int b, int64_t gb, int64_t c, int64_t n1, int tail,
                                          bool from_true_state, const double2 *st, int lane, float *partial, bool first) {
    CT s1[KB], s2[KB], hx[KB][2], hy[KB][2];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const double2 s = st[k * 32 + lane];
        s1[k] = static_cast<CT>(s.x);
        s2[k] = static_cast<CT>(s.y);
        const int64_t o = ((gb * KB + k) * g.C + c) * 2;
        hx[k][0] = hx[k][1] = hy[k][0] = hy[k][1] = CT(0);
        if (from_true_state) {  // consulted only when fewer than two samples are filtered (then S == 1)
            hx[k][0] = static_cast<CT>(g.state_x[o]);
            hx[k][1] = static_cast<CT>(g.state_x[o + 1]);
            hy[k][0] = static_cast<CT>(g.state_y[o]);
            hy[k][1] = static_cast<CT>(g.state_y[o + 1]);
        }
    }
    for (int e = 0; e < tail; ++e) {
        const int64_t n = n1 - tail + e;
        CT v = static_cast<CT>(g.x[c * g.ldx + n]);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const CT bb0 = static_cast<CT>(cd.b0[b][k]), bb1 = static_cast<CT>(cd.b1[b][k]), bb2 = static_cast<CT>(cd.b2[b][k]);
            const CT na1 = static_cast<CT>(-cd.a1[b][k]), na2 = static_cast<CT>(-cd.a2[b][k]);
            const CT y = fma_rn(bb0, v, s1[k]);
            s1[k] = fma_rn(na1, y, fma_rn(bb1, v, s2[k]));
            s2[k] = fma_rn(na2, y, bb2 * v);
            hx[k][1] = hx[k][0];
            hx[k][0] = v;
            hy[k][1] = hy[k][0];
            hy[k][0] = y;
            v = y;
        }
        if (partial == nullptr)
            g.y[gb * g.ldb + c * g.ldy + n] = static_cast<float>(v);
        else if (first)  // SUM mode: this warp's partial sum of the tail samples, [e][lane]
            partial[e * 32 + lane] = static_cast<float>(v);
        else
            partial[e * 32 + lane] += static_cast<float>(v);
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int64_t o = ((gb * KB + k) * g.C + c) * 2;
        g.state_x[o] = static_cast<double>(hx[k][0]);
        g.state_x[o + 1] = static_cast<double>(hx[k][1]);
        g.state_y[o] = static_cast<double>(hy[k][0]);
        g.state_y[o + 1] = static_cast<double>(hy[k][1]);
    }
}

template <int KB>
__global__ void __launch_bounds__(kMaxWarps * 32, 2)
bank_stack_kernel(const __grid_constant__ StackCoef<KB> cd, const __grid_constant__ StackGeom g) {
    // plain pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nthreads = g.W * 32;
    unsigned char *base_sm = smem_raw;
    if ((smem_u32(smem_raw) & 127u) != 0u) __trap();  // the swizzle needs 128-byte aligned tiles
    unsigned char *xring = base_sm;
    unsigned char *x64 = base_sm + kStagesX * kTileBytes;  // [32 x kCH] float64, same swizzle (2 * kNV columns)
    unsigned char *otile = base_sm + (kStagesX + 2) * kTileBytes + warp * warp_bytes(g.bpw, KB);
    double2 *stsm = reinterpret_cast<double2 *>(otile + kTileBytes);  // [slot * KB + k][lane]

    int lo[8];  // this lane's swizzled column offsets (see band_chunk)
#pragma unroll
    for (int jc = 0; jc < 8; ++jc) lo[jc] = lane * 128 + ((jc ^ (lane & 7)) << 4);

    // ---- the item: channel group x time segment (CTA-uniform) -----------------------------------
    const bool warm_pass = g.warm > 0;
    int64_t item = blockIdx.x;
    const int b_first = static_cast<int>(item % g.nsplit) * g.bps;  // this CTA's bands: [b_first, b_first + nb_cta)
    item /= g.nsplit;
    const int nb_cta = min(g.bps, g.n_bands - b_first);
    int64_t grp, j, n0, n1;
    if (warm_pass) {
        const int64_t sm1 = g.S - 1;
        grp = item / sm1;
        j = item - grp * sm1 + 1;
        n1 = j * g.Lseg;
        n0 = max(n1 - g.warm, static_cast<int64_t>(0));
    } else {
        grp = item / g.S;
        j = item - grp * g.S;
        n0 = j * g.Lseg;
        n1 = min(g.T, n0 + g.Lseg);
    }
    const int64_t c0 = grp * 32;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(32), g.C - c0));
    const int64_t c = c0 + lane;
    const bool live = lane < nrows;
    const bool do_tail = !warm_pass && (j == g.S - 1) && g.state_x != nullptr;
    const int tail = do_tail ? static_cast<int>(min(static_cast<int64_t>(2), n1 - n0)) : 0;
    const int64_t len = n1 - n0 - tail;
    const int64_t nch = (len + kCH - 1) / kCH;

    // ---- start states of my bands -> shared memory ------------------------------------------------
    for (int slot = 0; slot < g.bpw; ++slot) {
        const int bl = warp + slot * g.W;
        if (bl >= nb_cta) break;
        const int b = b_first + bl;
        const int64_t gb = g.band_id[b];
        // absolute sample at which band b starts in this launch: the segment start, or (warm-up) its own window
        const int64_t start_b = warm_pass ? max(n1 - static_cast<int64_t>(g.warm_b[b]), static_cast<int64_t>(0)) : n0;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            double a = 0.0, d = 0.0;
            if (live) {
                if (start_b == 0) {
                    if (g.state_x != nullptr) {
                        const int64_t o = ((gb * KB + k) * g.C + c) * 2;
                        const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                        const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                        a = cd.b1[b][k] * x1 + cd.b2[b][k] * x2 - cd.a1[b][k] * y1 - cd.a2[b][k] * y2;
                        d = cd.b2[b][k] * x1 - cd.a2[b][k] * y1;
                    }
                } else if (!warm_pass) {
                    const double *wsp = g.ws + (c * g.S + j);
                    a = wsp[((b * KB + k) * 2) * g.ws_stride];
                    d = wsp[((b * KB + k) * 2 + 1) * g.ws_stride];
                }
            }
            stsm[(slot * KB + k) * 32 + lane] = make_double2(a, d);
        }
    }

    auto issue_load = [&](int64_t i, int stage) {
        unsigned char *tile = xring + stage * kTileBytes;
        const int64_t base = i * kCH;
        const int cnt = static_cast<int>(min(len - base, static_cast<int64_t>(kCH)));
        if (cnt == kCH && g.vec_ok) {
            for (int idx = tid; idx < 32 * kNV; idx += nthreads) {
                const int r = idx / kNV, piece = idx % kNV;
                if (r < nrows) cp_async<16>(tile + col_offset(r, piece), g.x + (c0 + r) * g.ldx + n0 + base + piece * 4);
            }
        } else {
            for (int idx = tid; idx < 32 * kCH; idx += nthreads) {
                const int r = idx / kCH, e = idx % kCH;
                if (r < nrows && e < cnt) cp_async<4>(tile + elem_offset(r, e), g.x + (c0 + r) * g.ldx + n0 + base + e);
            }
        }
    };

#pragma unroll
    for (int st = 0; st < kStagesX; ++st) {
        if (st < nch) issue_load(st, st);
        cp_async_commit();
    }

    constexpr int RPI = 32 / kNV;  // rows per cooperative store instruction
    const int piece = lane % kNV;
    const int half = lane / kNV;
    int stage = 0;
    for (int64_t i = 0; i < nch; ++i) {
        cp_async_wait<kStagesX - 1>();
        __syncthreads();  // everyone's share of tile i has landed (also orders the state init above)
        const unsigned char *tile = xring + stage * kTileBytes;
        const int64_t base = i * kCH;
        const int cnt = static_cast<int>(min(len - base, static_cast<int64_t>(kCH)));

        if (((g.f64_mask >> b_first) & ((nb_cta >= 32 ? 0u : (1u << nb_cta)) - 1u)) != 0u) {  // CTA-uniform: float64 copy of the tile for this CTA's float64 bands
            for (int idx = tid; idx < 32 * kNV; idx += nthreads) {
                const int r = idx / kNV, p = idx % kNV;
                const float4 a = *reinterpret_cast<const float4 *>(tile + col_offset(r, p));
                *reinterpret_cast<double2 *>(x64 + col_offset(r, 2 * p)) = make_double2(static_cast<double>(a.x), static_cast<double>(a.y));
                *reinterpret_cast<double2 *>(x64 + col_offset(r, 2 * p + 1)) = make_double2(static_cast<double>(a.z), static_cast<double>(a.w));
            }
            __syncthreads();
        }

        for (int slot = 0; slot < g.bpw; ++slot) {
            const int bl = warp + slot * g.W;
            if (bl >= nb_cta) break;
            const int b = b_first + bl;
            double2 *st = stsm + slot * KB * 32;
            const bool is64 = (g.f64_mask >> b) & 1u;
            // ---- two bands at once where no per-band output tile is needed (warm-up, SUM) ---------------
            if (warm_pass || g.sum) {
                // ---- float32 bands: four (or two) at once on packed FFMA2 -------------------------------------------
                if (!is64 && KB <= 2) {
                    int nq = 1;  // consecutive float32 slots of this warp whose (warm-up) windows have begun
                    while (nq < 4 && slot + nq < g.bpw && bl + nq * g.W < nb_cta && (((g.f64_mask >> (b + nq * g.W)) & 1u) == 0u)) ++nq;
                    if (warm_pass) {
                        int ok = 0;
                        while (ok < nq && n0 + base >= max(n1 - static_cast<int64_t>(g.warm_b[b + ok * g.W]), static_cast<int64_t>(0))) ++ok;
                        nq = ok;
                    }
                    if (nq >= 2) {
                        const int take = nq >= 4 ? 4 : 2;
                        if (live) {
                            if (take == 4) {
                                const int bands[4] = {b, b + g.W, b + 2 * g.W, b + 3 * g.W};
                                double2 *const sts[4] = {st, st + KB * 32, st + 2 * KB * 32, st + 3 * KB * 32};
                                if (warm_pass)
                                    band_multi_p<KB, 2, 0>(cd, bands, tile, nullptr, sts, lane, cnt, lo);
                                else if (slot == 0)
                                    band_multi_p<KB, 2, 1>(cd, bands, tile, otile, sts, lane, cnt, lo);
                                else
                                    band_multi_p<KB, 2, 2>(cd, bands, tile, otile, sts, lane, cnt, lo);
                            } else {
                                const int bands[2] = {b, b + g.W};
                                double2 *const sts[2] = {st, st + KB * 32};
                                if (warm_pass)
                                    band_multi_p<KB, 1, 0>(cd, bands, tile, nullptr, sts, lane, cnt, lo);
                                else if (slot == 0)
                                    band_multi_p<KB, 1, 1>(cd, bands, tile, otile, sts, lane, cnt, lo);
                                else
                                    band_multi_p<KB, 1, 2>(cd, bands, tile, otile, sts, lane, cnt, lo);
                            }
                        }
                        slot += take - 1;  // the partner slots are done
                        continue;
                    }
                }
                const int bn = b + g.W;
                constexpr bool kPair64 = KB <= 2;  // two float64 bands of 3+ sections do not fit the register budget
                bool pair = slot + 1 < g.bpw && bl + g.W < nb_cta && (((g.f64_mask >> bn) & 1u) != 0u) == is64 && (kPair64 || !is64);
                if (pair && warm_pass) {  // both windows must have begun
                    const int64_t sa = max(n1 - static_cast<int64_t>(g.warm_b[b]), static_cast<int64_t>(0));
                    const int64_t sb = max(n1 - static_cast<int64_t>(g.warm_b[bn]), static_cast<int64_t>(0));
                    pair = n0 + base >= sa && n0 + base >= sb;
                }
                if (pair) {
                    double2 *stn = st + KB * 32;
                    if (live) {
                        if (warm_pass) {
                            if (is64) {
                                if constexpr (kPair64) band_pair<double, KB, 0>(cd, b, bn, x64, nullptr, st, stn, lane, cnt, lo);
                            } else
                                band_pair<float, KB, 0>(cd, b, bn, tile, nullptr, st, stn, lane, cnt, lo);
                        } else if (slot == 0) {
                            if (is64) {
                                if constexpr (kPair64) band_pair<double, KB, 1>(cd, b, bn, x64, otile, st, stn, lane, cnt, lo);
                            } else
                                band_pair<float, KB, 1>(cd, b, bn, tile, otile, st, stn, lane, cnt, lo);
                        } else {
                            if (is64) {
                                if constexpr (kPair64) band_pair<double, KB, 2>(cd, b, bn, x64, otile, st, stn, lane, cnt, lo);
                            } else
                                band_pair<float, KB, 2>(cd, b, bn, tile, otile, st, stn, lane, cnt, lo);
                        }
                    }
                    ++slot;  // the partner slot is done
                    continue;
                }
            }
            if (warm_pass) {
                const int64_t start_b = max(n1 - static_cast<int64_t>(g.warm_b[b]), static_cast<int64_t>(0));
                if (n0 + base < start_b) continue;  // this band's window has not begun yet
                if (live) {
                    if (is64)
                        band_chunk64<KB, 0>(cd, b, x64, nullptr, st, lane, cnt, lo);
                    else
                        band_chunk<float, KB, 0>(cd, b, tile, nullptr, st, lane, cnt, lo);
                }
                continue;
            }
            if (g.sum) {  // accumulate this warp's bands in its tile; reduced over the warps below
                if (live) {
                    if (slot == 0) {
                        if (is64)
                            band_chunk64<KB, 1>(cd, b, x64, otile, st, lane, cnt, lo);
                        else
                            band_chunk<float, KB, 1>(cd, b, tile, otile, st, lane, cnt, lo);
                    } else {
                        if (is64)
                            band_chunk64<KB, 2>(cd, b, x64, otile, st, lane, cnt, lo);
                        else
                            band_chunk<float, KB, 2>(cd, b, tile, otile, st, lane, cnt, lo);
                    }
                }
                continue;
            }
            if (live) {
                if (is64)
                    band_chunk64<KB, 1>(cd, b, x64, otile, st, lane, cnt, lo);
                else
                    band_chunk<float, KB, 1>(cd, b, tile, otile, st, lane, cnt, lo);
            }
            __syncwarp();
            // ---- the band's [32 x 64] tile leaves as coalesced rows of y[gb] ---------------------
            float *yb = g.y + static_cast<int64_t>(g.band_id[b]) * g.ldb + n0 + base;
            if (cnt == kCH && g.vec_ok) {
                float *yrow = yb + (c0 + half) * g.ldy + piece * 4;
                const int64_t ldy2 = RPI * g.ldy;
                if (nrows == 32) {
#pragma unroll 16
                    for (int t = 0; t < 32 / RPI; ++t)
                        store_row16(yrow + t * ldy2, *reinterpret_cast<const float4 *>(otile + col_offset(RPI * t + half, piece)));
                } else {
#pragma unroll 1
                    for (int t = 0; t < 32 / RPI; ++t)
                        if (RPI * t + half < nrows)
                            store_row16(yrow + t * ldy2, *reinterpret_cast<const float4 *>(otile + col_offset(RPI * t + half, piece)));
                }
            } else {
#pragma unroll 1
                for (int r = 0; r < nrows; ++r) {
                    float *dst = yb + (c0 + r) * g.ldy;
                    for (int e = lane; e < cnt; e += 32) dst[e] = *reinterpret_cast<const float *>(otile + elem_offset(r, e));
                }
            }
            __syncwarp();
        }
        if (g.sum && !warm_pass) {
            // ---- SUM: add the warps' partial tiles in warp order; the CTA stores ONE tile of y ----------
            __syncthreads();
            const unsigned char *t0 = base_sm + (kStagesX + 2) * kTileBytes;
            const int wb = warp_bytes(g.bpw, KB);
            float *yt = g.y + n0 + base;
            if (cnt == kCH && g.vec_ok) {
                for (int idx = tid; idx < 32 * kNV; idx += nthreads) {
                    const int r = idx / kNV, p = idx % kNV;
                    if (r >= nrows) continue;
                    const int off = col_offset(r, p);
                    float4 acc = *reinterpret_cast<const float4 *>(t0 + off);
                    for (int w = 1; w < g.W; ++w) {
                        const float4 a = *reinterpret_cast<const float4 *>(t0 + w * wb + off);
                        acc.x += a.x;
                        acc.y += a.y;
                        acc.z += a.z;
                        acc.w += a.w;
                    }
                    store_row16(yt + (c0 + r) * g.ldy + p * 4, acc);
                }
            } else {
                for (int idx = tid; idx < 32 * kCH; idx += nthreads) {
                    const int r = idx / kCH, e = idx % kCH;
                    if (r >= nrows || e >= cnt) continue;
                    const int off = elem_offset(r, e);
                    float acc = *reinterpret_cast<const float *>(t0 + off);
                    for (int w = 1; w < g.W; ++w) acc += *reinterpret_cast<const float *>(t0 + w * wb + off);
                    yt[(c0 + r) * g.ldy + e] = acc;
                }
            }
        }
        __syncthreads();  // tile i is free: refill its stage
        if (i + kStagesX < nch) issue_load(i + kStagesX, stage);
        cp_async_commit();
        stage = (stage + 1 == kStagesX) ? 0 : stage + 1;
    }
    cp_async_wait<0>();

    if (warm_pass) {
        if (live) {
            for (int slot = 0; slot < g.bpw; ++slot) {
                const int bl = warp + slot * g.W;
                if (bl >= nb_cta) break;
                const int b = b_first + bl;
                double *wsp = g.ws + (c * g.S + j);
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    const double2 s = stsm[(slot * KB + k) * 32 + lane];
                    wsp[((b * KB + k) * 2) * g.ws_stride] = s.x;
                    wsp[((b * KB + k) * 2 + 1) * g.ws_stride] = s.y;
                }
            }
        }
        return;
    }
    if (do_tail) {  // CTA-uniform
        const bool from_true_state = n0 == 0;
        float *partial = g.sum ? reinterpret_cast<float *>(otile) : nullptr;  // the output tile is free now
        if (live) {
            for (int slot = 0; slot < g.bpw; ++slot) {
                const int bl = warp + slot * g.W;
                if (bl >= nb_cta) break;
                const int b = b_first + bl;
                const double2 *st = stsm + slot * KB * 32;
                if ((g.f64_mask >> b) & 1u)
                    band_tail<double, KB>(cd, g, b, g.band_id[b], c, n1, tail, from_true_state, st, lane, partial, slot == 0);
                else
                    band_tail<float, KB>(cd, g, b, g.band_id[b], c, n1, tail, from_true_state, st, lane, partial, slot == 0);
            }
        }
        if (g.sum) {
            __syncthreads();
            if (warp == 0 && live) {
                const unsigned char *t0 = base_sm + (kStagesX + 2) * kTileBytes;
                const int wb = warp_bytes(g.bpw, KB);
                for (int e = 0; e < tail; ++e) {
                    float acc = reinterpret_cast<const float *>(t0)[e * 32 + lane];
                    for (int w = 1; w < g.W; ++w) acc += reinterpret_cast<const float *>(t0 + w * wb)[e * 32 + lane];
                    g.y[c * g.ldy + n1 - tail + e] = acc;
                }
            }
        }
    }
}

int ctas_per_sm(int W, int bpw, int KB) {
    const int n = kSmemPerSm / (cta_bytes(W, bpw, KB) + 1024);
    const int by_threads = 2048 / (W * 32);
    return std::max(1, std::min(std::min(n, by_threads), kMaxCtas));
}

// ------------------------------------------------------------------------------------------------------------
// Time-parallel warm-up for single-section bands (KB == 1: LogFilterBank, sums of biquads).
//
// The warm-up launch above gives a band's whole decay window to ONE lane-row per channel: the 20 Hz band of a
// LogFilterBank(32) needs 31 k samples, i.e. one float64 dependency chain of 31 k steps per (channel group, segment)
// while the other warps of the CTA have long finished -- 1.08 ms of the 3.6 ms the 32-channel shard of config 5 takes.
// A biquad is linear, so the window can be cut into 32 pieces that run in parallel from a ZERO state: with the free-
// response matrix A = [[-a1, 1], [-a2, 0]] of the DF2T state (s1, s2),
//     state(n1) = sum_p A^(n1 - end_p) z_p,        z_p = zero-state end state of piece p.
// One warp per (channel, band, segment): lane p filters piece p = [n1 - (p + 1) Lp, n1 - p Lp) of the channel
// (16-byte loads of its own stretch of the row; each 128-byte line is fetched once and then served by L1), raises A to
// p * Lp by repeated squaring in float64, and a fixed-order shuffle tree adds the 32 contributions (deterministic).
// The piece that contains sample 0 starts from the caller's DF1 state instead of zero, as the serial warm-up does.
// Critical path: Lp = warm / 32 samples instead of warm.
// ------------------------------------------------------------------------------------------------------------
struct M22 {
    double a, b, c, d;  // [[a, b], [c, d]]
};
__device__ __forceinline__ M22 mmul(const M22 &x, const M22 &y) {
    return M22{__fma_rn(x.a, y.a, x.b * y.c), __fma_rn(x.a, y.b, x.b * y.d), __fma_rn(x.c, y.a, x.d * y.c), __fma_rn(x.c, y.b, x.d * y.d)};
}
constexpr int kWarm1Warps = 4;
constexpr int kW1Samples = 32;                 // samples of every piece staged per step: 128-byte rows
constexpr int kW1TileBytes = 32 * kW1Samples * 4;
// Piece length for a window of W samples: an odd multiple of 32 samples (a whole number of staged tiles), >= W / 32.
__host__ __device__ inline int64_t warm1_piece(int64_t W) {
    int64_t m = (W + 32 * 32 - 1) / (32 * 32);
    if (m < 1) m = 1;
    if ((m & 1) == 0) ++m;
    return m * 32;
}
// The warp stages tile i of all 32 pieces together -- [32 rows x 128 bytes], rows one piece apart in the channel's row --
// with coalesced 16-byte cp.async (a warp instruction covers 4 rows x 128 bytes: 4 L1/L2 line requests), double buffered,
// XOR-swizzled so that lane p reads "its" row with conflict-free 128-bit shared loads.  (A first version let every lane
// load its own stretch straight from global memory: 32 line requests per warp instruction, and ncu showed the kernel
// waiting on those loads -- long_scoreboard -- with the FP64 pipe 36 % busy.)
template <typename CT>
__device__ __forceinline__ void warm1_tile(const unsigned char *tile, int lane, CT b0, CT b1, CT b2, CT na1, CT na2, CT &s1, CT &s2) {
    auto step = [&](float xf) {
        const CT v = static_cast<CT>(xf);
        const CT y = fma_rn(b0, v, s1);
        s1 = fma_rn(na1, y, fma_rn(b1, v, s2));
        s2 = fma_rn(na2, y, b2 * v);
    };
    const unsigned char *row = tile + lane * 128;
    float4 a = *reinterpret_cast<const float4 *>(row + ((0 ^ (lane & 7)) << 4));
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int vn = (v + 1) & 7;
        const float4 nxt = *reinterpret_cast<const float4 *>(row + ((vn ^ (lane & 7)) << 4));
        step(a.x);
        step(a.y);
        step(a.z);
        step(a.w);
        a = nxt;
    }
}

__global__ void __launch_bounds__(kWarm1Warps * 32, 7) bank_warm1_kernel(const __grid_constant__ StackCoef<1> cd, const __grid_constant__ StackGeom g,
                                                                        int64_t nitems) {
    __shared__ __align__(128) unsigned char w1_smem[kWarm1Warps][2][kW1TileBytes];
    const int lane = threadIdx.x & 31;
    const int64_t item = static_cast<int64_t>(blockIdx.x) * kWarm1Warps + (threadIdx.x >> 5);
    if (item >= nitems) return;  // warp-uniform (no CTA barriers below)
    unsigned char *buf = &w1_smem[threadIdx.x >> 5][0][0];
    // item = (segment j - 1, band b, channel c), c fastest: the warps of a CTA run the SAME band on neighbouring channels, so
    // they finish together (with b fastest a CTA held a 992-sample and three short pieces: ncu showed 36 % of the warp slots
    // active); the bands of one segment follow each other, so their overlapping windows are served by L2.
    const int64_t c = item % g.C;
    const int64_t r = item / g.C;
    const int b = static_cast<int>(r % g.n_bands);
    const int64_t j = r / g.n_bands + 1;
    const int64_t n1 = j * g.Lseg;
    const int64_t W = g.warm_b[b];
    const int64_t Lp = warm1_piece(W);                   // piece length, samples
    const int nl = static_cast<int>((W + Lp - 1) / Lp);  // pieces (= active lanes) that cover the window, <= 32
    const int64_t ntiles = Lp / kW1Samples;
    const int64_t hi = n1 - lane * Lp, lo = hi - Lp;     // my piece; samples before 0 do not exist
    const bool empty = lane >= nl || hi <= 0;
    const double b0 = cd.b0[b][0], b1 = cd.b1[b][0], b2 = cd.b2[b][0], a1 = cd.a1[b][0], a2 = cd.a2[b][0];
    const bool is64 = (g.f64_mask >> b) & 1u;
    const float *xrow = g.x + c * g.ldx;

    auto issue = [&](int64_t i, int stage) {
        unsigned char *tile = buf + stage * kW1TileBytes;
        if (g.vec_ok) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int rr = (lane >> 3) + 4 * t, col = lane & 7;
                const int64_t n = n1 - static_cast<int64_t>(rr + 1) * Lp + i * kW1Samples;  // first sample of tile i of piece rr
                if (rr < nl && n >= 0) cp_async<16>(tile + rr * 128 + ((col ^ (rr & 7)) << 4), xrow + n + col * 4);
            }
        } else {
            for (int idx = lane; idx < 32 * kW1Samples; idx += 32) {
                const int rr = idx / kW1Samples, e = idx % kW1Samples;
                const int64_t n = n1 - static_cast<int64_t>(rr + 1) * Lp + i * kW1Samples;
                if (rr < nl && n >= 0) cp_async<4>(tile + rr * 128 + (((e >> 2) ^ (rr & 7)) << 4) + (e & 3) * 4, xrow + n + e);
            }
        }
    };

    double z1 = 0.0, z2 = 0.0;
    float f1 = 0.f, f2 = 0.f;
    issue(0, 0);
    cp_async_commit();
    for (int64_t i = 0; i < ntiles; ++i) {
        if (i + 1 < ntiles) issue(i + 1, static_cast<int>((i + 1) & 1));
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const int64_t n = lo + i * kW1Samples;  // first sample of my tile
        if (!empty && n >= 0) {
            if (n == 0 && g.state_x != nullptr) {  // my piece contains sample 0: start from the caller's DF1 state, not from zero
                const int64_t o = (static_cast<int64_t>(g.band_id[b]) * g.C + c) * 2;
                const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                z1 = b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
                z2 = b2 * x1 - a2 * y1;
                f1 = static_cast<float>(z1);
                f2 = static_cast<float>(z2);
            }
            const unsigned char *tile = buf + static_cast<int>(i & 1) * kW1TileBytes;
            if (is64)
                warm1_tile<double>(tile, lane, b0, b1, b2, -a1, -a2, z1, z2);
            else
                warm1_tile<float>(tile, lane, static_cast<float>(b0), static_cast<float>(b1), static_cast<float>(b2), static_cast<float>(-a1),
                                  static_cast<float>(-a2), f1, f2);
        }
        __syncwarp();  // everyone is done with this stage before the copy of tile i + 2 lands in it
    }
    if (!is64) {
        z1 = static_cast<double>(f1);
        z2 = static_cast<double>(f2);
    }
    // A^(lane * Lp): M = A^Lp by binary exponentiation (warp-uniform), then M^lane from M, M^2, M^4, M^8, M^16
    M22 M{1.0, 0.0, 0.0, 1.0}, Q{-a1, 1.0, -a2, 0.0};
    for (int64_t e = Lp; e > 0; e >>= 1) {
        if (e & 1) M = mmul(M, Q);
        Q = mmul(Q, Q);
    }
    M22 P{1.0, 0.0, 0.0, 1.0};
    for (int bit = 0; bit < 5 && ((max(nl, 1) - 1) >> bit) != 0; ++bit) {  // warp-uniform trip count
        if ((lane >> bit) & 1) P = mmul(P, M);
        M = mmul(M, M);
    }
    double t1 = empty ? 0.0 : __fma_rn(P.a, z1, P.b * z2);
    double t2 = empty ? 0.0 : __fma_rn(P.c, z1, P.d * z2);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        t1 += __shfl_xor_sync(0xffffffffu, t1, off);
        t2 += __shfl_xor_sync(0xffffffffu, t2, off);
    }
    if (lane == 0) {
        double *wsp = g.ws + (c * g.S + j);
        wsp[(b * 2) * g.ws_stride] = t1;
        wsp[(b * 2 + 1) * g.ws_stride] = t2;
    }
}

template <int KB>
int launch_stack_kb(const SosSection *sec, StackGeom g, const Segmentation &seg, cudaStream_t stream) {
    StackCoef<KB> cd;
    for (int b = 0; b < 32; ++b)
        for (int k = 0; k < KB; ++k) {
            const SosSection id{1.0, 0.0, 0.0, 0.0, 0.0};
            const SosSection &s = b < g.n_bands ? sec[b * KB + k] : id;
            cd.b0[b][k] = s.b0;
            cd.b1[b][k] = s.b1;
            cd.b2[b][k] = s.b2;
            cd.a1[b][k] = s.a1;
            cd.a2[b][k] = s.a2;
        }
    auto kern = bank_stack_kernel<KB>;
    TFX_ENSURE_SMEM(kern, cta_bytes(kMaxWarps, 32 / kMaxWarps, KB));
    const int smem = cta_bytes(g.W, g.bpw, KB);
    const int64_t G = (g.C + 31) / 32;
    if (seg.S > 1) {
        const bool serial_warm = std::getenv("TFX_BS_SERIAL_WARM") != nullptr;  // developer A/B switch (read per call: tests toggle it)
        // Serial warm-up: one wave of CTAs, its time is the longest band's window (a single dependency chain).  Time-
        // parallel warm-up: (nearly) the same arithmetic spread over all SMs, its time is the total work.  Both sustain
        // ~0.9 G lane-samples per ms on a B200 (32 x 32 x 60 s: 1.08 ms serial, 0.5 ms parallel; 32 x 256 x 60 s: 1.6 ms
        // against 2.7 ms), the serial chain runs ~35 ns per sample: take the parallel kernel while
        // total / longest < ~30 000 concurrent lane-streams.
        bool parallel_warm = false;
        if constexpr (KB == 1) {
            int64_t total = 0, longest = 1;
            for (int b = 0; b < g.n_bands; ++b) {
                const int64_t W = g.warm_b[b], Lp = warm1_piece(W);
                total += (W + Lp - 1) / Lp * Lp;
                longest = std::max<int64_t>(longest, W);
            }
            const char *force = std::getenv("TFX_BS_PARALLEL_WARM");  // developer A/B switch
            // (SUM banks run a single band split, so a warp's serial chain is the sum of its bands' windows: twice the room;
            //  32 biquads x 256 ch x 60 s: 9.25 ms serial, 8.80 ms parallel)
            parallel_warm = !serial_warm && (force != nullptr || static_cast<double>(g.C) * (seg.S - 1) * total / longest < (g.sum ? 60000.0 : 30000.0));
        }
        if constexpr (KB == 1) {
            if (parallel_warm) {
                const int64_t nitems = g.C * (seg.S - 1) * g.n_bands;
                TFX_REQUIRE((nitems + kWarm1Warps - 1) / kWarm1Warps < (int64_t(1) << 31), "filterbank: too many warm-up items for one launch");
                bank_warm1_kernel<<<static_cast<unsigned>((nitems + kWarm1Warps - 1) / kWarm1Warps), kWarm1Warps * 32, 0, stream>>>(cd, g, nitems);
                TFX_CHECK_LAUNCH("bank_warm1_kernel");
            }
        }
        if (!parallel_warm) {
            StackGeom gw = g;
            gw.warm = seg.warm;
            kern<<<static_cast<unsigned>(G * (seg.S - 1) * g.nsplit), g.W * 32, smem, stream>>>(cd, gw);
            TFX_CHECK_LAUNCH("bank_stack_kernel(warm-up)");
        }
    }
    g.warm = 0;
    kern<<<static_cast<unsigned>(G * seg.S * g.nsplit), g.W * 32, smem, stream>>>(cd, g);
    TFX_CHECK_LAUNCH("bank_stack_kernel");
    return TFX_OK;
}


// How one launch is cut: time segments (sos_plan.cpp) and, for STACK banks over few channels / short signals, the
// bands over up to 4 CTAs per (channel group, segment) so that the grid still fills the GPU (x is re-read per split:
// 4/N of the output traffic each).  items = CTAs of the main launch, resident = CTAs the GPU holds at once.
struct StackPlan {
    int nsplit, bps, W, bpw;
    Segmentation seg;
    int64_t items, resident;
};
StackPlan plan_stack(int64_t C, int64_t T, int nb, int Kb, int64_t warm_max, bool sum, bool no_split) {
    const int64_t G = (C + 31) / 32, lanes = G * 32;
    StackPlan best{};
    for (int ns = 1; ns <= (sum ? 1 : 4) && ns <= nb; ns *= 2) {
        StackPlan p{};
        p.bps = (nb + ns - 1) / ns;
        p.nsplit = (nb + p.bps - 1) / p.bps;
        p.W = std::min(kMaxWarps, p.bps);
        p.bpw = (p.bps + p.W - 1) / p.W;
        p.resident = static_cast<int64_t>(sm_count()) * ctas_per_sm(p.W, p.bpw, Kb);
        const int64_t s_cap = std::max<int64_t>(p.resident / (G * p.nsplit), 1);  // one wave: more CTAs only add warm-up work (1.73 / 2.73 waves measured 15 / 7 % slower on 256 ch x 60 s)
        p.seg = choose_segmentation(lanes, T, warm_max, s_cap * lanes, no_split);
        p.items = G * p.seg.S * p.nsplit;
        if (ns == 1 || p.items > best.items) best = p;
        if (best.items * 10 >= best.resident * 9) break;  // (nearly) a full wave of CTAs: enough
    }
    return best;
}

}  // namespace

bool bank_stack_worthwhile(int64_t C, int64_t T, int nb, int Kb, int64_t warm_max, bool sum, bool no_split) {
    const int64_t wa = warm_max < 0 ? 0 : (warm_max + kCH - 1) / kCH * kCH;
    const StackPlan p = plan_stack(C, T, nb, Kb, wa, sum, no_split || warm_max < 0);
    return p.items * 4 >= p.resident;  // below a quarter of a wave the band-per-lane kernel fills the GPU better
}

bool bank_stack_tile_ok(int N, int Kb, int64_t C) {
    const int64_t G = (C + 31) / 32;
    return N >= 1 && Kb >= 1 && Kb <= 4 && C * 5 >= G * 32 * 4 && G * 64 < (int64_t(1) << 31);
}

int launch_bank_stack(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t ldb, const SosSection *sec,
                      const int *band_id, const int64_t *warm_b, uint32_t f64_mask, int nb, int Kb, bool sum, bool no_split,
                      void *workspace, size_t workspace_bytes, double *state_x, double *state_y, cudaStream_t stream) {
    // Band splits (STACK over few channels): a CTA takes the local bands [s * bps, (s + 1) * bps).  In the caller's order
    // (ascending frequency for a LogFilterBank) split 0 would hold all the long-warm-up, float64 bands and the grid --
    // a single wave of CTAs -- would wait for it: the 32-channel shard of config 5 ran 2.87 ms with SM active cycles
    // between 2.9 M and 5.5 M (profiles/r2_ncu_summary.md).  Deal the bands to the splits round-robin in order of cost
    // instead (float64 first, then longer warm-up); band_id keeps the output planes and state blocks where they belong.
    std::vector<SosSection> sec_p;
    int band_id_p[32];
    int64_t warm_p[32];
    {
        int64_t wm = 0;
        bool ns_local = no_split;
        for (int b = 0; b < nb; ++b) {
            if (warm_b[b] < 0) ns_local = true;
            wm = std::max<int64_t>(wm, warm_b[b] < 0 ? 0 : (warm_b[b] + kCH - 1) / kCH * kCH);
        }
        if (wm > (int64_t(1) << 30)) ns_local = true;
        const StackPlan pl0 = plan_stack(C, T, nb, Kb, wm, sum, ns_local);
        if (!sum && pl0.nsplit > 1) {
            std::vector<int> order(nb);
            for (int b = 0; b < nb; ++b) order[b] = b;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                const int fa = (f64_mask >> a) & 1u, fb = (f64_mask >> b) & 1u;
                if (fa != fb) return fa > fb;
                return warm_b[a] > warm_b[b];
            });
            int filled[4] = {0, 0, 0, 0}, cap[4] = {0, 0, 0, 0};
            for (int sp = 0; sp < pl0.nsplit; ++sp) cap[sp] = std::max(0, std::min(pl0.bps, nb - sp * pl0.bps));
            int perm[32];  // perm[new local position] = caller's band
            int sp = 0;
            for (int t = 0; t < nb; ++t) {
                while (filled[sp] >= cap[sp]) sp = (sp + 1) % pl0.nsplit;
                perm[sp * pl0.bps + filled[sp]++] = order[t];
                sp = (sp + 1) % pl0.nsplit;
            }
            sec_p.resize(static_cast<size_t>(nb) * Kb);
            uint32_t mask_p = 0;
            for (int q = 0; q < nb; ++q) {
                const int o = perm[q];
                for (int k = 0; k < Kb; ++k) sec_p[static_cast<size_t>(q) * Kb + k] = sec[static_cast<size_t>(o) * Kb + k];
                band_id_p[q] = band_id[o];
                warm_p[q] = warm_b[o];
                mask_p |= ((f64_mask >> o) & 1u) << q;
            }
            sec = sec_p.data();
            band_id = band_id_p;
            warm_b = warm_p;
            f64_mask = mask_p;
        }
    }
    StackGeom g{};
    g.sum = sum ? 1 : 0;
    g.x = x;
    g.y = y;
    g.ldx = ldx;
    g.ldy = ldy;
    g.ldb = sum ? 0 : ldb;
    g.C = C;
    g.T = T;
    g.n_bands = nb;
    g.f64_mask = f64_mask;
    g.state_x = state_x;
    g.state_y = state_y;
    int64_t warm_max = 0;
    for (int b = 0; b < 32; ++b) {
        g.band_id[b] = band_id[b < nb ? b : 0];
        const int64_t w = b < nb ? warm_b[b] : 0;
        if (w < 0) no_split = true;
        const int64_t wa = w < 0 ? 0 : (w + kCH - 1) / kCH * kCH;  // band windows start on chunk boundaries
        g.warm_b[b] = static_cast<int>(std::min<int64_t>(wa, int64_t(1) << 30));
        warm_max = std::max(warm_max, wa);
    }
    if (warm_max > (int64_t(1) << 30)) no_split = true;
    const StackPlan pl = plan_stack(C, T, nb, Kb, warm_max, sum, no_split);
    g.nsplit = pl.nsplit;
    g.bps = pl.bps;
    g.W = pl.W;
    g.bpw = pl.bpw;
    const Segmentation seg = pl.seg;
    g.S = seg.S;
    g.Lseg = seg.Lseg;
    g.ws = static_cast<double *>(workspace);
    g.ws_stride = C * seg.S;
    if (seg.S > 1) {
        const size_t need = static_cast<size_t>(2 * Kb * nb) * static_cast<size_t>(C * seg.S) * 8;
        if (workspace == nullptr || workspace_bytes < need) {
            set_error("filterbank: workspace of %zu bytes needed, %zu given (query tfx_filterbank_workspace_bytes)", need, workspace_bytes);
            return TFX_EWORKSPACE;
        }
    }
    const size_t esz = 4;
    g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && ((ldx * esz) % 16 == 0) &&
               ((ldy * esz) % 16 == 0) && ((g.ldb * esz) % 16 == 0);
    switch (Kb) {
        case 1: return launch_stack_kb<1>(sec, g, seg, stream);
        case 2: return launch_stack_kb<2>(sec, g, seg, stream);
        case 3: return launch_stack_kb<3>(sec, g, seg, stream);
        case 4: return launch_stack_kb<4>(sec, g, seg, stream);
        default: set_error("filterbank: Kb must be in [1, 4]"); return TFX_EINVAL;
    }
}

int64_t bank_stack_max_streams() { return static_cast<int64_t>(sm_count()) * kMaxCtas * 32; }

}  // namespace tfx
