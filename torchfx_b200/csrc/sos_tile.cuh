// sos_tile.cuh -- the channel-tile kernel template (see sos_tile.cu for the description), shared by
// sos_tile.cu (series cascades) and bank_tile.cu (parallel `+` banks).  Everything lives in an
// anonymous namespace: each translation unit instantiates only what it launches.
#pragma once
#include <algorithm>
#include <cstdint>
#include <type_traits>

#include "common.cuh"
#include "sos_kernels.h"
#include "stream_common.cuh"

namespace tfx {
namespace {

#ifndef TFX_T_STAGES
#define TFX_T_STAGES 2
#endif
#ifndef TFX_T_WARPS
#define TFX_T_WARPS 1
#endif
#ifndef TFX_T_PREFETCH
#define TFX_T_PREFETCH 1
#endif
constexpr int kStages = TFX_T_STAGES;
constexpr int kWarps = TFX_T_WARPS;
constexpr int kTileBytes = 32 * 256;
constexpr int kWarpSmem = kStages * kTileBytes;
constexpr int kCtaSmem = kWarps * kWarpSmem + 128;  // + slack to align the tiles to 128 B
constexpr int kCtasPerSm = std::min(32, kSmemPerSm / (kCtaSmem + 1024));
constexpr int kWarpsPerSm = kCtasPerSm * kWarps;
static_assert(kCtasPerSm >= 1, "CTA does not fit in shared memory");

struct TileGeom {
    const void *x;
    void *y;
    int64_t ldx, ldy, C, T;
    int64_t S, Lseg, warm;
    int64_t nitems;  // G * S (main) or G * (S - 1) (warm-up)
    int64_t G;       // channel groups of 32
    void *ws;        // [2K][C * S]
    int64_t ws_stride;
    double *state_x;
    double *state_y;
    unsigned long long *counter;  // NULL: item = global warp id
    int vec_ok;
    unsigned f64_mask;  // MixedF only: sections that run the float64 recurrence
};

// byte offset of the 16-byte column v (0..15) of row r inside a swizzled tile
__device__ __forceinline__ int col_offset(int r, int v) { return (v >> 3) * 4096 + r * 128 + (((v & 7) ^ (r & 7)) << 4); }
template <typename IO>
__device__ __forceinline__ int elem_offset(int r, int e) {
    constexpr int EPV = 16 / sizeof(IO);
    return col_offset(r, e / EPV) + (e % EPV) * static_cast<int>(sizeof(IO));
}

// Compute-type tag: float32 recurrence except for the sections flagged in TileGeom::f64_mask,
// which run in float64 (float32 signal in and out of each such section).  Lets a chain like
// LoButterworth | ParametricEQ | HiShelving pay for float64 only in the one section whose
// float32 round-off would break the 1e-5 bar (sos_plan.cpp picks the mask with a probe).
template <unsigned MASK>
struct MixedF {};  // MASK is a compile-time constant: a run-time per-section branch costs more than float64 everywhere

template <typename CT>
struct CtTraits {
    using Coef = CT;   // element type of the SosCoef kernel parameter
    using Store = CT;  // element type of the workspace / tracked history
    static constexpr bool heavy = sizeof(CT) == 8;
};
template <unsigned MASK>
struct CtTraits<MixedF<MASK>> {
    using Coef = float;
    using Store = double;
    static constexpr bool heavy = false;
};

// DF2T state of one stream + the recurrence, uniform precision
template <typename CT, int K>
struct Cascade {
    CT s1[K], s2[K];
    __device__ __forceinline__ void set(int k, double a, double b) {
        s1[k] = static_cast<CT>(a);
        s2[k] = static_cast<CT>(b);
    }
    __device__ __forceinline__ CT get1(int k) const { return s1[k]; }
    __device__ __forceinline__ CT get2(int k) const { return s2[k]; }
    template <typename IO>
    __device__ __forceinline__ IO step(const SosCoef<CT, K> &cf, const SosCoefD<K> &, unsigned, IO x) {
        return static_cast<IO>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(x)));
    }
    // one sample, recording every section's input / output (DF1 history)
    template <typename IO>
    __device__ __forceinline__ IO step_tracked(const SosCoef<CT, K> &cf, const SosCoefD<K> &, unsigned, IO x, CT (&hx)[K][2],
                                               CT (&hy)[K][2]) {
        CT v = static_cast<CT>(x);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const CT y = fma_rn(cf.b0[k], v, s1[k]);
            s1[k] = fma_rn(cf.na1[k], y, fma_rn(cf.b1[k], v, s2[k]));
            s2[k] = fma_rn(cf.na2[k], y, cf.b2[k] * v);
            hx[k][1] = hx[k][0];
            hx[k][0] = v;
            hy[k][1] = hy[k][0];
            hy[k][0] = y;
            v = y;
        }
        return static_cast<IO>(v);
    }
};

// mixed precision: per-section float32 or float64 state, float32 signal between sections
template <unsigned MASK, int K>
struct Cascade<MixedF<MASK>, K> {
    float f1[K], f2[K];
    double d1[K], d2[K];
    __device__ __forceinline__ void set(int k, double a, double b) {
        f1[k] = static_cast<float>(a);
        f2[k] = static_cast<float>(b);
        d1[k] = a;
        d2[k] = b;
    }
    __device__ __forceinline__ double get1(int k) const { return d1[k]; }
    __device__ __forceinline__ double get2(int k) const { return d2[k]; }
    __device__ __forceinline__ void sync_views() {  // keep get1/get2 meaningful for f32 sections
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (!((MASK >> k) & 1u)) {
                d1[k] = f1[k];
                d2[k] = f2[k];
            }
    }
    __device__ __forceinline__ float section(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, unsigned, int k, float v,
                                             double &xin, double &yout) {
        if ((MASK >> k) & 1u) {  // folds at compile time once the section loop is unrolled
            const double vd = static_cast<double>(v);
            const double y = __fma_rn(cd.b0[k], vd, d1[k]);
            d1[k] = __fma_rn(-cd.a1[k], y, __fma_rn(cd.b1[k], vd, d2[k]));
            d2[k] = __fma_rn(-cd.a2[k], y, cd.b2[k] * vd);
            xin = vd;
            yout = y;
            return static_cast<float>(y);
        }
        const float y = __fmaf_rn(cf.b0[k], v, f1[k]);
        f1[k] = __fmaf_rn(cf.na1[k], y, __fmaf_rn(cf.b1[k], v, f2[k]));
        f2[k] = __fmaf_rn(cf.na2[k], y, cf.b2[k] * v);
        xin = v;
        yout = y;
        return y;
    }
    template <typename IO>
    __device__ __forceinline__ IO step(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, unsigned mask, IO x) {
        float v = static_cast<float>(x);
        double a, b;
#pragma unroll
        for (int k = 0; k < K; ++k) v = section(cf, cd, mask, k, v, a, b);
        return static_cast<IO>(v);
    }
    template <typename IO>
    __device__ __forceinline__ IO step_tracked(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, unsigned mask, IO x,
                                               double (&hx)[K][2], double (&hy)[K][2]) {
        float v = static_cast<float>(x);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double a, b;
            v = section(cf, cd, mask, k, v, a, b);
            hx[k][1] = hx[k][0];
            hx[k][0] = a;
            hy[k][1] = hy[k][0];
            hy[k][0] = b;
        }
        return static_cast<IO>(v);
    }
};

// Parallel topology (`f1 + f2 + ...`, filter/__base.py:1019-1026): the K sections are K / KB
// branches of KB sections in series, every branch fed by the same x; the branch outputs are
// rounded to the I/O type and added in branch order, exactly the reference's
// `out = zeros_like(x); out += f(x)` loop.  Section index k = branch * KB + section, which is
// also the layout of the bank's DF1 state [N, Kb, C, 2].  The branches are independent
// dependency chains, so this topology has K / KB-way instruction-level parallelism per lane.
template <typename CT, int KB>
struct Par {};
template <typename CT, int KB>
struct CtTraits<Par<CT, KB>> {
    using Coef = CT;
    using Store = CT;
    static constexpr bool heavy = sizeof(CT) == 8;
};
template <typename CT, int KB, int K>
struct Cascade<Par<CT, KB>, K> {
    static_assert(K % KB == 0, "parallel bank: K must be a whole number of branches");
    CT s1[K], s2[K];
    __device__ __forceinline__ void set(int k, double a, double b) {
        s1[k] = static_cast<CT>(a);
        s2[k] = static_cast<CT>(b);
    }
    __device__ __forceinline__ CT get1(int k) const { return s1[k]; }
    __device__ __forceinline__ CT get2(int k) const { return s2[k]; }
    __device__ __forceinline__ void sync_views() {}
    template <typename IO>
    __device__ __forceinline__ IO step(const SosCoef<CT, K> &cf, const SosCoefD<K> &, unsigned, IO x) {
        const CT xin = static_cast<CT>(x);
        CT v = xin;
        IO acc = IO(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k % KB == 0) v = xin;
            const CT y = fma_rn(cf.b0[k], v, s1[k]);
            s1[k] = fma_rn(cf.na1[k], y, fma_rn(cf.b1[k], v, s2[k]));
            s2[k] = fma_rn(cf.na2[k], y, cf.b2[k] * v);
            v = y;
            if (k % KB == KB - 1) acc = (k == KB - 1) ? static_cast<IO>(v) : acc + static_cast<IO>(v);  // 0 + y0 == y0
        }
        return acc;
    }
    template <typename IO>
    __device__ __forceinline__ IO step_tracked(const SosCoef<CT, K> &cf, const SosCoefD<K> &, unsigned, IO x, CT (&hx)[K][2],
                                               CT (&hy)[K][2]) {
        const CT xin = static_cast<CT>(x);
        CT v = xin;
        IO acc = IO(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k % KB == 0) v = xin;
            const CT y = fma_rn(cf.b0[k], v, s1[k]);
            s1[k] = fma_rn(cf.na1[k], y, fma_rn(cf.b1[k], v, s2[k]));
            s2[k] = fma_rn(cf.na2[k], y, cf.b2[k] * v);
            hx[k][1] = hx[k][0];
            hx[k][0] = v;
            hy[k][1] = hy[k][0];
            hy[k][0] = y;
            v = y;
            if (k % KB == KB - 1) acc = (k == KB - 1) ? static_cast<IO>(v) : acc + static_cast<IO>(v);  // 0 + y0 == y0
        }
        return acc;
    }
};

// Parallel topology, mixed precision: the first N64 branches run the float64 recurrence, the others
// float32 (TFX_PREC_AUTO when the float32-hostile branches are a prefix of the bank, the usual case of a
// bank listed by rising frequency).  x is widened once per sample for all float64 branches; each
// float64 branch rounds its output to float32 before the branch-ordered sum, like the reference.
template <int N64, int KB>
struct ParMixed {};
template <int N64, int KB>
struct CtTraits<ParMixed<N64, KB>> {
    using Coef = float;
    using Store = double;
    static constexpr bool heavy = N64 * KB >= 4;  // register budget: mostly-float64 banks run at half the residency
};
template <int N64, int KB, int K>
struct Cascade<ParMixed<N64, KB>, K> {
    static_assert(K % KB == 0 && N64 >= 1 && N64 * KB < K, "mixed parallel bank: 1 <= N64 < branches");
    static constexpr int K64 = N64 * KB;  // sections [0, K64) are float64
    float f1[K], f2[K];
    double d1[K64], d2[K64];
    __device__ __forceinline__ void set(int k, double a, double b) {
        f1[k] = static_cast<float>(a);
        f2[k] = static_cast<float>(b);
        if (k < K64) {
            d1[k] = a;
            d2[k] = b;
        }
    }
    __device__ __forceinline__ double get1(int k) const { return k < K64 ? d1[k] : static_cast<double>(f1[k]); }
    __device__ __forceinline__ double get2(int k) const { return k < K64 ? d2[k] : static_cast<double>(f2[k]); }
    __device__ __forceinline__ void sync_views() {}
    template <bool TRACK>
    __device__ __forceinline__ float run(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, float x, double (*hx)[2], double (*hy)[2]) {
        const double xd = static_cast<double>(x);
        float acc = 0.f;
        double vd = xd;
        float vf = x;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < K64) {
                if (k % KB == 0) vd = xd;
                const double y = __fma_rn(cd.b0[k], vd, d1[k]);
                d1[k] = __fma_rn(-cd.a1[k], y, __fma_rn(cd.b1[k], vd, d2[k]));
                d2[k] = __fma_rn(-cd.a2[k], y, cd.b2[k] * vd);
                if (TRACK) {
                    hx[k][1] = hx[k][0];
                    hx[k][0] = vd;
                    hy[k][1] = hy[k][0];
                    hy[k][0] = y;
                }
                vd = y;
                if (k % KB == KB - 1) acc += static_cast<float>(y);
            } else {
                if (k % KB == 0) vf = x;
                const float y = __fmaf_rn(cf.b0[k], vf, f1[k]);
                f1[k] = __fmaf_rn(cf.na1[k], y, __fmaf_rn(cf.b1[k], vf, f2[k]));
                f2[k] = __fmaf_rn(cf.na2[k], y, cf.b2[k] * vf);
                if (TRACK) {
                    hx[k][1] = hx[k][0];
                    hx[k][0] = vf;
                    hy[k][1] = hy[k][0];
                    hy[k][0] = y;
                }
                vf = y;
                if (k % KB == KB - 1) acc += y;
            }
        }
        return acc;
    }
    template <typename IO>
    __device__ __forceinline__ IO step(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, unsigned, IO x) {
        return static_cast<IO>(run<false>(cf, cd, static_cast<float>(x), nullptr, nullptr));
    }
    template <typename IO>
    __device__ __forceinline__ IO step_tracked(const SosCoef<float, K> &cf, const SosCoefD<K> &cd, unsigned, IO x, double (&hx)[K][2],
                                               double (&hy)[K][2]) {
        return static_cast<IO>(run<true>(cf, cd, static_cast<float>(x), hx, hy));
    }
};

// Register budget: float64-heavy instantiations with more than TFX_T_HEAVY_K sections run at half the residency
// (up to 255 registers); the others are held to 65536 / (13 x 32) registers so that 13 warps per SM stay resident.
#ifndef TFX_T_HEAVY_K
#define TFX_T_HEAVY_K 4
#endif
template <typename IO, typename CT, int K>
__global__ void __launch_bounds__(kWarps * 32, (CtTraits<CT>::heavy && K > TFX_T_HEAVY_K) ? (kCtasPerSm + 1) / 2 : kCtasPerSm)
sos_tile_kernel(const __grid_constant__ SosCoef<typename CtTraits<CT>::Coef, K> cf, const __grid_constant__ SosCoefD<K> cd,
                const __grid_constant__ TileGeom g) {
    using Store = typename CtTraits<CT>::Store;
    using Tr = IoTraits<IO>;
    using Vec = typename Tr::Vec;
    constexpr int VEC = Tr::VEC;
    constexpr int CH = 256 / sizeof(IO);  // samples per chunk
    constexpr int UV = K <= 2 ? 16 : (K <= 4 ? 8 : 4);

    // Plain pointer arithmetic on the __shared__ array (no integer round trip): the compiler keeps the
    // address space and emits LDS / STS instead of generic LD / ST for every tile access.
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *ring = smem_raw + warp * kWarpSmem;
    if ((smem_u32(smem_raw) & 127u) != 0u) __trap();  // the swizzle needs 128-byte aligned tiles
    const IO *__restrict__ xg = static_cast<const IO *>(g.x);
    IO *__restrict__ yg = static_cast<IO *>(g.y);
    const bool warm_pass = g.warm > 0;
    const int piece = lane & 15;
    const int half = lane >> 4;

    int64_t item = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
    for (;;) {
        if (g.counter != nullptr) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(g.counter, 1ULL);
            item = static_cast<int64_t>(__shfl_sync(0xffffffffu, t, 0));
        }
        if (item >= g.nitems) break;

        // ---- the item: channel group x time segment (all warp-uniform) -----------------------
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
        const bool from_true_state = n0 == 0;
        const bool do_tail = !warm_pass && (j == g.S - 1) && g.state_x != nullptr;
        const int64_t len = n1 - n0;
        const int64_t nch = (len + CH - 1) / CH;

        // ---- start state (DF2T) ----------------------------------------------------------------
        Cascade<CT, K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) st.set(k, 0.0, 0.0);
        if (live) {
            if (from_true_state) {
                if (g.state_x != nullptr) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
                        const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                        const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                        st.set(k, cd.b1[k] * x1 + cd.b2[k] * x2 - cd.a1[k] * y1 - cd.a2[k] * y2, cd.b2[k] * x1 - cd.a2[k] * y1);
                    }
                }
            } else if (!warm_pass) {
                const Store *wsp = static_cast<const Store *>(g.ws) + (c * g.S + j);
#pragma unroll
                for (int k = 0; k < K; ++k) st.set(k, wsp[(2 * k) * g.ws_stride], wsp[(2 * k + 1) * g.ws_stride]);
            }
        }

        // row (2t + half), 16-byte piece `piece`: this lane's share of every cooperative copy
        const IO *xrow = xg + (c0 + half) * g.ldx + n0 + piece * VEC;
        IO *yrow = yg + (c0 + half) * g.ldy + n0 + piece * VEC;
        const int64_t ldx2 = 2 * g.ldx, ldy2 = 2 * g.ldy;

        auto issue_load = [&](int64_t i, int stage) {
            unsigned char *tile = ring + stage * kTileBytes;
            const int64_t base = i * CH;
            const int cnt = static_cast<int>(min(len - base, static_cast<int64_t>(CH)));
            if (cnt == CH && g.vec_ok) {
                if (nrows == 32) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) cp_async<16>(tile + col_offset(2 * t + half, piece), xrow + t * ldx2 + base);
                } else {
#pragma unroll 1
                    for (int t = 0; t < 16; ++t)
                        if (2 * t + half < nrows) cp_async<16>(tile + col_offset(2 * t + half, piece), xrow + t * ldx2 + base);
                }
            } else {
#pragma unroll 1
                for (int r = 0; r < nrows; ++r) {
                    const IO *src = xg + (c0 + r) * g.ldx + n0 + base;
                    for (int e = lane; e < cnt; e += 32) cp_async<sizeof(IO)>(tile + elem_offset<IO>(r, e), src + e);
                }
            }
        };

#pragma unroll
        for (int st = 0; st < kStages; ++st) {
            if (st < nch) issue_load(st, st);
            cp_async_commit();
        }

        // DF1 history of every section, only maintained over the channel's last two chunks
        Store hx[K][2], hy[K][2];
        if (do_tail) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                hx[k][0] = hx[k][1] = hy[k][0] = hy[k][1] = Store(0);
                if (live && from_true_state) {  // consulted only when fewer than two samples are filtered (then S == 1)
                    const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
                    hx[k][0] = static_cast<Store>(g.state_x[o]);
                    hx[k][1] = static_cast<Store>(g.state_x[o + 1]);
                    hy[k][0] = static_cast<Store>(g.state_y[o]);
                    hy[k][1] = static_cast<Store>(g.state_y[o + 1]);
                }
            }
        }

        int stage = 0;
        for (int64_t i = 0; i < nch; ++i) {
            cp_async_wait<kStages - 1>();
            __syncwarp();
            unsigned char *tile = ring + stage * kTileBytes;
            const int64_t base = i * CH;
            const int cnt = static_cast<int>(min(len - base, static_cast<int64_t>(CH)));
            const bool tracked = do_tail && i >= nch - 2;

            // ---- filter my row (channel) in place ----------------------------------------------
            if (live) {
                if (cnt == CH && !tracked) {
                    // the next vector is fetched before the current one is filtered and stored (the
                    // compiler cannot hoist a load over the in-place store on its own)
                    Vec a = *reinterpret_cast<const Vec *>(tile + col_offset(lane, 0));
#pragma unroll UV
                    for (int v = 0; v < 16; ++v) {
#if TFX_T_PREFETCH
                        const Vec nxt = *reinterpret_cast<const Vec *>(tile + col_offset(lane, (v + 1) & 15));
#else
                        if (v > 0) a = *reinterpret_cast<const Vec *>(tile + col_offset(lane, v));
#endif
                        a.x = st.step(cf, cd, g.f64_mask, a.x);
                        a.y = st.step(cf, cd, g.f64_mask, a.y);
                        if constexpr (VEC == 4) {
                            a.z = st.step(cf, cd, g.f64_mask, a.z);
                            a.w = st.step(cf, cd, g.f64_mask, a.w);
                        }
                        *reinterpret_cast<Vec *>(tile + col_offset(lane, v)) = a;
#if TFX_T_PREFETCH
                        a = nxt;
#endif
                    }
                } else if (!tracked) {
                    for (int e = 0; e < cnt; ++e) {
                        IO *p = reinterpret_cast<IO *>(tile + elem_offset<IO>(lane, e));
                        *p = st.step(cf, cd, g.f64_mask, *p);
                    }
                } else {
                    for (int e = 0; e < cnt; ++e) {
                        IO *p = reinterpret_cast<IO *>(tile + elem_offset<IO>(lane, e));
                        *p = st.step_tracked(cf, cd, g.f64_mask, *p, hx, hy);
                    }
                }
            }
            __syncwarp();

            // ---- write the tile back, coalesced ------------------------------------------------
            if (!warm_pass) {
                if (cnt == CH && g.vec_ok) {
                    if (nrows == 32) {
#pragma unroll
                        for (int t = 0; t < 16; ++t)
                            st_stream16(yrow + t * ldy2 + base, *reinterpret_cast<const Vec *>(tile + col_offset(2 * t + half, piece)));
                    } else {
#pragma unroll 1
                        for (int t = 0; t < 16; ++t)
                            if (2 * t + half < nrows)
                                st_stream16(yrow + t * ldy2 + base, *reinterpret_cast<const Vec *>(tile + col_offset(2 * t + half, piece)));
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < nrows; ++r) {
                        IO *dst = yg + (c0 + r) * g.ldy + n0 + base;
                        for (int e = lane; e < cnt; e += 32) dst[e] = *reinterpret_cast<const IO *>(tile + elem_offset<IO>(r, e));
                    }
                }
            }
            __syncwarp();

            if (i + kStages < nch) issue_load(i + kStages, stage);
            cp_async_commit();
            stage = (stage + 1 == kStages) ? 0 : stage + 1;
        }
        cp_async_wait<0>();

        if (live) {
            if (warm_pass) {
                if constexpr (!std::is_floating_point<CT>::value) st.sync_views();
                Store *wsp = static_cast<Store *>(g.ws) + (c * g.S + j);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    wsp[(2 * k) * g.ws_stride] = st.get1(k);
                    wsp[(2 * k + 1) * g.ws_stride] = st.get2(k);
                }
            } else if (do_tail) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
                    g.state_x[o] = static_cast<double>(hx[k][0]);
                    g.state_x[o + 1] = static_cast<double>(hx[k][1]);
                    g.state_y[o] = static_cast<double>(hy[k][0]);
                    g.state_y[o + 1] = static_cast<double>(hy[k][1]);
                }
            }
        }
        if (g.counter == nullptr) break;
        __syncwarp();
    }
}

template <typename IO, typename CT, int K>
int launch_tile_k(const SosSection *sec, TileGeom g, const Segmentation &seg, unsigned long long *counter, cudaStream_t stream) {
    using CoefT = typename CtTraits<CT>::Coef;
    SosCoef<CoefT, K> cf;
    SosCoefD<K> cd;
    for (int k = 0; k < K; ++k) {
        cf.b0[k] = static_cast<CoefT>(sec[k].b0);
        cf.b1[k] = static_cast<CoefT>(sec[k].b1);
        cf.b2[k] = static_cast<CoefT>(sec[k].b2);
        cf.na1[k] = static_cast<CoefT>(-sec[k].a1);
        cf.na2[k] = static_cast<CoefT>(-sec[k].a2);
        cd.b0[k] = sec[k].b0;
        cd.b1[k] = sec[k].b1;
        cd.b2[k] = sec[k].b2;
        cd.a1[k] = sec[k].a1;
        cd.a2[k] = sec[k].a2;
    }
    auto kern = sos_tile_kernel<IO, CT, K>;
    TFX_ENSURE_SMEM(kern, kCtaSmem);
    if (seg.S > 1) {
        TileGeom gw = g;
        gw.warm = seg.warm;
        gw.nitems = g.G * (seg.S - 1);
        gw.counter = nullptr;
        const int64_t grid = (gw.nitems + kWarps - 1) / kWarps;
        kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, gw);
        TFX_CHECK_LAUNCH("sos_tile_kernel(warm-up)");
    }
    g.warm = 0;
    g.nitems = g.G * seg.S;
    g.counter = nullptr;
    int64_t grid = (g.nitems + kWarps - 1) / kWarps;
    const int64_t resident = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    if (seg.S > 1 && grid > resident && counter != nullptr) {
        g.counter = counter;
        TFX_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
        grid = resident;
    }
    kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, g);
    TFX_CHECK_LAUNCH("sos_tile_kernel");
    return TFX_OK;
}

}  // namespace
}  // namespace tfx
