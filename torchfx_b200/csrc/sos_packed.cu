// sos_packed.cu -- float32 cascade kernel on Blackwell's packed FP32 pipe (FFMA2).
//
// Same streaming design as sos_cascade.cu (stream = channel x time segment, cp.async tiles
// through padded shared-memory rows, persistent warps pulling work items), but every THREAD
// owns TWO streams (rows l and l+32 of a 64-row tile) and runs both recurrences in the two
// halves of 64-bit register pairs with fma.rn.f32x2 / mul.rn.f32x2 (SASS FFMA2 / FMUL2, new
// on sm_100): the 5*K FMA-class operations per sample of the DF2T cascade cost 5*K issue
// slots per sample PAIR.  The first ncu captures of the scalar kernel (profiles/) showed it
// bounded by instruction issue and dependency stalls (issue-active ~55 %, FMA pipe ~42 %),
// not by HBM; halving the issue slots moves the bound to the memory system.
//
// Only the steady state is packed: ragged chunk ends, the two-sample DF1 tail and the state
// hand-over run the scalar recurrence on the individual halves (.x = stream A, .y = stream B).
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "sos_kernels.h"
#include "stream_common.cuh"

namespace tfx {
namespace {

#ifndef TFX_P_WARPS
#define TFX_P_WARPS 2
#endif
#ifndef TFX_P_STAGES
#define TFX_P_STAGES 2
#endif
constexpr int kWarps = TFX_P_WARPS;
constexpr int kStages = TFX_P_STAGES;
constexpr int kRows = 64;  // streams per warp
constexpr int kTableBytes = 3 * kRows * 8;
constexpr int kWarpSmem = kStages * kRows * kPitch + kTableBytes;
constexpr int kCtaSmem = kWarps * kWarpSmem;
constexpr int kCtasPerSm = kSmemPerSm / (kCtaSmem + 1024) < 16 ? kSmemPerSm / (kCtaSmem + 1024) : 16;
constexpr int kWarpsPerSm = kCtasPerSm * kWarps;
static_assert(kCtasPerSm >= 1, "CTA does not fit in shared memory");
static_assert(kRowBytes == 256, "packed kernel assumes 256-byte rows");

// A packed pair lives in ONE 64-bit register for its whole life (pack at the tile load,
// unpack at the tile store): going through float2 makes ptxas shuffle halves with MOVs
// around every operation, which costs as many issue slots as FFMA2 saves.
using p2 = unsigned long long;

__device__ __forceinline__ p2 pack2(float lo, float hi) {
    p2 r;
    // volatile: ptxas otherwise re-materialises the pair (2 MOVs) at every use of it
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(p2 v) { return __uint_as_float(static_cast<unsigned>(v)); }
__device__ __forceinline__ float hi2(p2 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) {
    p2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ p2 mul2(p2 a, p2 b) {
    p2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int K>
struct PairCoef {
    p2 b0[K], b1[K], b2[K], na1[K], na2[K];  // both halves hold the same coefficient
    float sb0[K], sb1[K], sb2[K], sna1[K], sna2[K];  // scalar copies for the ragged paths
    p2 one;                                          // {1.0f, 1.0f}
};

template <int K>
__device__ __forceinline__ p2 pair_step(const PairCoef<K> &cf, p2 (&s1)[K], p2 (&s2)[K], p2 v) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const p2 y = fma2(cf.b0[k], v, s1[k]);
        s1[k] = fma2(cf.na1[k], y, fma2(cf.b1[k], v, s2[k]));
        s2[k] = fma2(cf.na2[k], y, mul2(cf.b2[k], v));
        v = y;
    }
    return v;
}

// scalar recurrence on one half (H = 0: low half / stream A, H = 1: high half / stream B)
template <int K, int H>
__device__ __forceinline__ float half_step(const PairCoef<K> &cf, p2 (&s1)[K], p2 (&s2)[K], float v) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t1 = H ? hi2(s1[k]) : lo2(s1[k]);
        float t2 = H ? hi2(s2[k]) : lo2(s2[k]);
        const float y = __fmaf_rn(cf.sb0[k], v, t1);
        t1 = __fmaf_rn(cf.sna1[k], y, __fmaf_rn(cf.sb1[k], v, t2));
        t2 = __fmaf_rn(cf.sna2[k], y, cf.sb2[k] * v);
        s1[k] = H ? pack2(lo2(s1[k]), t1) : pack2(t1, hi2(s1[k]));
        s2[k] = H ? pack2(lo2(s2[k]), t2) : pack2(t2, hi2(s2[k]));
        v = y;
    }
    return v;
}

template <int K>
__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm)
sos_pair_kernel(const __grid_constant__ PairCoef<K> cf, const __grid_constant__ SosCoefD<K> cd, const __grid_constant__ Geom g) {
    constexpr int CH = 64;  // samples per chunk (256-byte rows of float)
    constexpr int UV = K <= 2 ? 16 : (K <= 4 ? 8 : 4);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wsm = smem_raw + warp * kWarpSmem;
    int64_t *t_offx = reinterpret_cast<int64_t *>(wsm + kStages * kRows * kPitch);
    int64_t *t_offy = t_offx + kRows;
    int64_t *t_len = t_offy + kRows;

    const float *__restrict__ xg = static_cast<const float *>(g.x);
    float *__restrict__ yg = static_cast<float *>(g.y);

    const bool warm_pass = g.warm > 0;
    const int64_t nitems = (g.nstreams + kRows - 1) / kRows;
    int64_t item = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
    const int piece = lane & 15;
    const int half = lane >> 4;

    for (;;) {
        if (g.counter != nullptr) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(g.counter, 1ULL);
            item = static_cast<int64_t>(__shfl_sync(0xffffffffu, t, 0));
        }
        if (item >= nitems) break;

        // ---- my two streams: h = 0 -> row lane, h = 1 -> row lane + 32 ----------------------
        int64_t c[2], j[2], n0[2], n1[2], len[2];
        bool live[2], from_true[2], do_tail[2];
        int tail[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t q = item * kRows + lane + 32 * h;
            live[h] = q < g.nstreams;
            c[h] = j[h] = n0[h] = n1[h] = 0;
            if (live[h]) {
                if (warm_pass) {
                    const int64_t sm1 = g.S - 1;
                    c[h] = q / sm1;
                    j[h] = q - c[h] * sm1 + 1;
                    n1[h] = j[h] * g.Lseg;
                    n0[h] = max(n1[h] - g.warm, static_cast<int64_t>(0));
                } else {
                    c[h] = q / g.S;
                    j[h] = q - c[h] * g.S;
                    n0[h] = j[h] * g.Lseg;
                    n1[h] = min(g.T, n0[h] + g.Lseg);
                }
            }
            from_true[h] = live[h] && n0[h] == 0;
            do_tail[h] = live[h] && !warm_pass && (j[h] == g.S - 1) && g.state_x != nullptr;
            tail[h] = do_tail[h] ? static_cast<int>(min(static_cast<int64_t>(2), n1[h] - n0[h])) : 0;
            len[h] = live[h] ? n1[h] - n0[h] - tail[h] : 0;
        }

        // ---- start state (DF2T), halves = streams ---------------------------------------------
        p2 s1[K], s2[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s1[k] = 0ull;
            s2[k] = 0ull;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (from_true[h]) {
                if (g.state_x != nullptr) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int64_t o = (static_cast<int64_t>(k) * g.C + c[h]) * 2;
                        const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                        const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                        const float a = static_cast<float>(cd.b1[k] * x1 + cd.b2[k] * x2 - cd.a1[k] * y1 - cd.a2[k] * y2);
                        const float b = static_cast<float>(cd.b2[k] * x1 - cd.a2[k] * y1);
                        s1[k] = h ? pack2(lo2(s1[k]), a) : pack2(a, hi2(s1[k]));
                        s2[k] = h ? pack2(lo2(s2[k]), b) : pack2(b, hi2(s2[k]));
                    }
                }
            } else if (live[h] && !warm_pass) {
                const float *wsp = static_cast<const float *>(g.ws) + (c[h] * g.S + j[h]);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float a = wsp[(2 * k) * g.ws_stride], b = wsp[(2 * k + 1) * g.ws_stride];
                    s1[k] = h ? pack2(lo2(s1[k]), a) : pack2(a, hi2(s1[k]));
                    s2[k] = h ? pack2(lo2(s2[k]), b) : pack2(b, hi2(s2[k]));
                }
            }
        }

#pragma unroll
        for (int h = 0; h < 2; ++h) {
            t_offx[lane + 32 * h] = c[h] * g.ldx + n0[h];
            t_offy[lane + 32 * h] = c[h] * g.ldy + n0[h];
            t_len[lane + 32 * h] = len[h];
        }
        int64_t maxlen = max(len[0], len[1]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
        const int64_t nch = (maxlen + CH - 1) / CH;
        __syncwarp();

        auto issue_load = [&](int64_t i, int stage) {
            unsigned char *buf = wsm + stage * (kRows * kPitch);
            const int64_t base = i * CH;
            const bool all_full = __all_sync(0xffffffffu, len[0] - base >= CH && len[1] - base >= CH);
            if (all_full && g.vec_ok) {
#pragma unroll
                for (int t = 0; t < kRows / 2; ++t) {
                    const int r = 2 * t + half;
                    cp_async<16>(buf + r * kPitch + piece * 16, xg + t_offx[r] + base + piece * 4);
                }
            } else {
#pragma unroll 1
                for (int t = 0; t < kRows / 2; ++t) {
                    const int r = 2 * t + half;
                    const int64_t rem = t_len[r] - base;
                    const float *src = xg + t_offx[r] + base + piece * 4;
                    unsigned char *dst = buf + r * kPitch + piece * 16;
                    if (g.vec_ok && rem >= (piece + 1) * 4) {
                        cp_async<16>(dst, src);
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            if (piece * 4 + v < rem) cp_async<4>(dst + v * 4, src + v);
                    }
                }
            }
        };

#pragma unroll
        for (int st = 0; st < kStages; ++st) {
            if (st < nch) issue_load(st, st);
            cp_async_commit();
        }

        int stage = 0;
        for (int64_t i = 0; i < nch; ++i) {
            cp_async_wait<kStages - 1>();
            __syncwarp();
            unsigned char *buf = wsm + stage * (kRows * kPitch);
            const int64_t base = i * CH;
            const bool mine_full = len[0] - base >= CH && len[1] - base >= CH;
            const bool all_full = __all_sync(0xffffffffu, mine_full);

            // ---- filter my two rows in place ----------------------------------------------------
            if (mine_full) {
                float4 *ra = reinterpret_cast<float4 *>(buf + lane * kPitch);
                float4 *rb = reinterpret_cast<float4 *>(buf + (lane + 32) * kPitch);
#pragma unroll UV
                for (int v = 0; v < 16; ++v) {
                    float4 a = ra[v], b = rb[v];
                    // x * 1 makes the freshly packed pair the RESULT of a packed op: ptxas then keeps
                    // one copy for its three uses instead of re-forming it (6 MOVs) per use.
                    const p2 y0 = pair_step<K>(cf, s1, s2, mul2(pack2(a.x, b.x), cf.one));
                    const p2 y1 = pair_step<K>(cf, s1, s2, mul2(pack2(a.y, b.y), cf.one));
                    const p2 y2 = pair_step<K>(cf, s1, s2, mul2(pack2(a.z, b.z), cf.one));
                    const p2 y3 = pair_step<K>(cf, s1, s2, mul2(pack2(a.w, b.w), cf.one));
                    ra[v] = make_float4(lo2(y0), lo2(y1), lo2(y2), lo2(y3));
                    rb[v] = make_float4(hi2(y0), hi2(y1), hi2(y2), hi2(y3));
                }
            } else {
                float *ra = reinterpret_cast<float *>(buf + lane * kPitch);
                float *rb = reinterpret_cast<float *>(buf + (lane + 32) * kPitch);
                const int ca = static_cast<int>(max(static_cast<int64_t>(0), min(len[0] - base, static_cast<int64_t>(CH))));
                const int cb = static_cast<int>(max(static_cast<int64_t>(0), min(len[1] - base, static_cast<int64_t>(CH))));
                for (int e = 0; e < ca; ++e) ra[e] = half_step<K, 0>(cf, s1, s2, ra[e]);
                for (int e = 0; e < cb; ++e) rb[e] = half_step<K, 1>(cf, s1, s2, rb[e]);
            }
            __syncwarp();

            // ---- write the 64 rows back, coalesced ------------------------------------------------
            if (!warm_pass) {
                if (all_full && g.vec_ok) {
#pragma unroll
                    for (int t = 0; t < kRows / 2; ++t) {
                        const int r = 2 * t + half;
                        const float4 v = *reinterpret_cast<const float4 *>(buf + r * kPitch + piece * 16);
                        st_stream16(yg + t_offy[r] + base + piece * 4, v);
                    }
                } else {
#pragma unroll 1
                    for (int t = 0; t < kRows / 2; ++t) {
                        const int r = 2 * t + half;
                        const int64_t rem = t_len[r] - base;
                        const unsigned char *src = buf + r * kPitch + piece * 16;
                        float *dst = yg + t_offy[r] + base + piece * 4;
                        if (g.vec_ok && rem >= (piece + 1) * 4) {
                            st_stream16(dst, *reinterpret_cast<const float4 *>(src));
                        } else {
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                if (piece * 4 + v < rem) dst[v] = reinterpret_cast<const float *>(src)[v];
                        }
                    }
                }
            }
            __syncwarp();

            if (i + kStages < nch) issue_load(i + kStages, stage);
            cp_async_commit();
            stage = (stage + 1 == kStages) ? 0 : stage + 1;
        }
        cp_async_wait<0>();

        // ---- epilogue per stream: warm-up state out, or last two samples + DF1 state out -------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (live[h] && warm_pass) {
                float *wsp = static_cast<float *>(g.ws) + (c[h] * g.S + j[h]);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    wsp[(2 * k) * g.ws_stride] = h ? hi2(s1[k]) : lo2(s1[k]);
                    wsp[(2 * k + 1) * g.ws_stride] = h ? hi2(s2[k]) : lo2(s2[k]);
                }
            }
            if (do_tail[h]) {
                float hx[K][2], hy[K][2];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int64_t o = (static_cast<int64_t>(k) * g.C + c[h]) * 2;
                    hx[k][0] = static_cast<float>(g.state_x[o]);
                    hx[k][1] = static_cast<float>(g.state_x[o + 1]);
                    hy[k][0] = static_cast<float>(g.state_y[o]);
                    hy[k][1] = static_cast<float>(g.state_y[o + 1]);
                }
                for (int e = 0; e < tail[h]; ++e) {
                    const int64_t n = n1[h] - tail[h] + e;
                    float v = xg[c[h] * g.ldx + n];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        float t1 = h ? hi2(s1[k]) : lo2(s1[k]);
                        float t2 = h ? hi2(s2[k]) : lo2(s2[k]);
                        const float y = __fmaf_rn(cf.sb0[k], v, t1);
                        t1 = __fmaf_rn(cf.sna1[k], y, __fmaf_rn(cf.sb1[k], v, t2));
                        t2 = __fmaf_rn(cf.sna2[k], y, cf.sb2[k] * v);
                        s1[k] = h ? pack2(lo2(s1[k]), t1) : pack2(t1, hi2(s1[k]));
                        s2[k] = h ? pack2(lo2(s2[k]), t2) : pack2(t2, hi2(s2[k]));
                        hx[k][1] = hx[k][0];
                        hx[k][0] = v;
                        hy[k][1] = hy[k][0];
                        hy[k][0] = y;
                        v = y;
                    }
                    yg[c[h] * g.ldy + n] = v;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int64_t o = (static_cast<int64_t>(k) * g.C + c[h]) * 2;
                    g.state_x[o] = static_cast<double>(hx[k][0]);
                    g.state_x[o + 1] = static_cast<double>(hx[k][1]);
                    g.state_y[o] = static_cast<double>(hy[k][0]);
                    g.state_y[o + 1] = static_cast<double>(hy[k][1]);
                }
            }
        }
        if (g.counter == nullptr) break;
        __syncwarp();  // the per-warp tables are rewritten by the next item
    }
}

template <int K>
int launch_pair_k(const SosSection *sec, Geom g, const Segmentation &seg, cudaStream_t stream) {
    PairCoef<K> cf;
    SosCoefD<K> cd;
    for (int k = 0; k < K; ++k) {
        const float b0 = static_cast<float>(sec[k].b0), b1 = static_cast<float>(sec[k].b1), b2 = static_cast<float>(sec[k].b2);
        const float na1 = static_cast<float>(-sec[k].a1), na2 = static_cast<float>(-sec[k].a2);
        auto dup = [](float v) {
            uint32_t u;
            memcpy(&u, &v, 4);
            return (static_cast<unsigned long long>(u) << 32) | u;
        };
        cf.b0[k] = dup(b0);
        cf.b1[k] = dup(b1);
        cf.b2[k] = dup(b2);
        cf.na1[k] = dup(na1);
        cf.na2[k] = dup(na2);
        cf.sb0[k] = b0;
        cf.sb1[k] = b1;
        cf.sb2[k] = b2;
        cf.sna1[k] = na1;
        cf.sna2[k] = na2;
        cf.one = dup(1.0f);
        cd.b0[k] = sec[k].b0;
        cd.b1[k] = sec[k].b1;
        cd.b2[k] = sec[k].b2;
        cd.a1[k] = sec[k].a1;
        cd.a2[k] = sec[k].a2;
    }
    auto kern = sos_pair_kernel<K>;
    TFX_ENSURE_SMEM(kern, kCtaSmem);
    const int64_t per_cta = static_cast<int64_t>(kWarps) * kRows;
    if (seg.S > 1) {
        Geom gw = g;
        gw.warm = seg.warm;
        gw.nstreams = g.C * (seg.S - 1);
        gw.counter = nullptr;
        const int64_t grid = (gw.nstreams + per_cta - 1) / per_cta;
        kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, gw);
        TFX_CHECK_LAUNCH("sos_pair_kernel(warm-up)");
    }
    g.warm = 0;
    g.nstreams = g.C * seg.S;
    g.counter = nullptr;
    int64_t grid = (g.nstreams + per_cta - 1) / per_cta;
    const int64_t resident = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    if (seg.S > 1 && grid > resident) {
        g.counter = static_cast<unsigned long long *>(g.ws_base);
        TFX_CUDA_TRY(cudaMemsetAsync(g.counter, 0, sizeof(unsigned long long), stream));
        grid = resident;
    }
    kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, g);
    TFX_CHECK_LAUNCH("sos_pair_kernel");
    return TFX_OK;
}

}  // namespace

int64_t packed_stream_capacity() { return static_cast<int64_t>(sm_count()) * kWarpsPerSm * kRows; }

int launch_packed_pass(const SosSection *sec, int k, Geom g, const Segmentation &seg, cudaStream_t stream) {
    switch (k) {
        case 1: return launch_pair_k<1>(sec, g, seg, stream);
        case 2: return launch_pair_k<2>(sec, g, seg, stream);
        case 3: return launch_pair_k<3>(sec, g, seg, stream);
        case 4: return launch_pair_k<4>(sec, g, seg, stream);
        case 5: return launch_pair_k<5>(sec, g, seg, stream);
        case 6: return launch_pair_k<6>(sec, g, seg, stream);
        case 7: return launch_pair_k<7>(sec, g, seg, stream);
        case 8: return launch_pair_k<8>(sec, g, seg, stream);
        default: set_error("internal: pass with %d sections", k); return TFX_EINVAL;
    }
}

}  // namespace tfx
