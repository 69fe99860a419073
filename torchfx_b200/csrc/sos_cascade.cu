// sos_cascade.cu -- the fused SOS / biquad cascade for sm_100a.
//
// Replaces (reference) cuda/biquad_forward.cu:49-92 + cuda/parallel_scan.cu:117-364, which
// run K x (forcing kernel + 3-phase f64 Blelloch scan) full passes over the signal
// (~184 B/sample of HBM traffic at K = 4).  Here the whole cascade is ONE pass:
// 8 B/sample (read x f32, write y f32), all K sections' state in registers.
//
// Decomposition ("streams").  Channel c is cut into S time segments of Lseg samples;
// stream q = c*S + j filters segment j sequentially.  One THREAD owns one stream, so the
// DF2T recurrence   y = b0*x + s1;  s1 = b1*x - a1*y + s2;  s2 = b2*x - a2*y
// is 5 FMA-class instructions per section per sample with coefficients read straight from
// the constant bank (kernel parameters) and no cross-lane traffic at all.
//
// Memory path.  A warp owns 32 streams.  Per pipeline step it moves one 256-byte chunk of
// each of its 32 streams HBM -> shared memory with 16 fully-coalesced cp.async.cg
// (LDGSTS.128) instructions: lanes 0-15 fetch one stream's chunk, lanes 16-31 the next
// stream's, so every request is 256 contiguous bytes.  Rows sit in shared memory with a
// 272-byte pitch so the per-lane 128-bit row reads/writes of the compute phase are
// bank-conflict free.  The thread filters its row IN PLACE in shared memory, then the warp
// writes the 32 rows back with coalesced 128-bit streaming stores.  Two chunks per warp
// are in flight (cp.async groups); warps never synchronise with each other
// (__syncwarp only), 12 warps / SM.
//
// Work distribution.  SMs do not stream at the same speed (L2-slice distance, DRAM page
// luck): with one equal share per warp the launch ends when the slowest SM does (13 % of
// SM-time idle in the first ncu capture, profiles/).  Long signals are therefore cut into
// ~8 segments per resident warp and one wave of persistent warps pulls 32-stream work
// items from a global counter.
//
// Segment start state.  Segment 0 starts from the caller's DF1 state (converted to DF2T in
// f64).  Segment j > 0 gets its state from a warm-up launch of the SAME kernel that runs
// the recurrence over the `warm` samples preceding the segment without storing output
// (sos_plan.cpp explains why that is exact to the working precision).  Two launches, so
// the scheme is also correct in place (y == x).
//
// Final state.  The last segment's thread handles its final two samples outside the
// pipelined loop while recording every section's input/output, which is exactly the
// reference's DF1 state {v[n-1], v[n-2]} (cpu/iir_cpu.cpp:125-130,150-155).
#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "sos_kernels.h"
#include "sos_plan.h"
#include "stream_common.cuh"

namespace tfx {
namespace {

#ifndef TFX_WARPS
#define TFX_WARPS 4
#endif
#ifndef TFX_STAGES
#define TFX_STAGES 2
#endif
constexpr int kWarps = TFX_WARPS;    // warps per CTA
constexpr int kStages = TFX_STAGES;  // chunks in flight per warp
constexpr int kTableBytes = 3 * 32 * 8;
constexpr int kWarpSmem = kStages * 32 * kPitch + kTableBytes;
constexpr int kCtaSmem = kWarps * kWarpSmem;
constexpr int kCtasPerSm = kSmemPerSm / (kCtaSmem + 1024) < 8 ? kSmemPerSm / (kCtaSmem + 1024) : 8;
constexpr int kWarpsPerSm = kCtasPerSm * kWarps;
static_assert(kCtasPerSm >= 1, "CTA does not fit in shared memory");

template <typename IO, typename CT, int K>
__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm)
sos_stream_kernel(const __grid_constant__ SosCoef<CT, K> cf, const __grid_constant__ SosCoefD<K> cd,
                  const __grid_constant__ Geom g) {
    using Tr = IoTraits<IO>;
    using Vec = typename Tr::Vec;
    constexpr int CH = Tr::CHUNK;
    constexpr int VEC = Tr::VEC;
    constexpr int NV = kNvec;  // 16-byte vectors per row
    constexpr int UV = K <= 2 ? 16 : (K <= 4 ? 8 : 4);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wsm = smem_raw + warp * kWarpSmem;
    int64_t *t_offx = reinterpret_cast<int64_t *>(wsm + kStages * 32 * kPitch);
    int64_t *t_offy = t_offx + 32;
    int64_t *t_len = t_offy + 32;

    const IO *__restrict__ xg = static_cast<const IO *>(g.x);
    IO *__restrict__ yg = static_cast<IO *>(g.y);

    // ---- work items: one item = 32 consecutive streams -----------------------------------
    // Static: item = global warp id (one item per warp).  Dynamic (g.counter != NULL): the
    // launch is one wave of persistent warps that pull items from a global counter, so SMs
    // that stream faster take more of the (many, short) segments and all finish together.
    const bool warm_pass = g.warm > 0;
    const int64_t nitems = (g.nstreams + 31) / 32;
    int64_t item = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  for (;;) {
    if (g.counter != nullptr) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(g.counter, 1ULL);
        item = static_cast<int64_t>(__shfl_sync(0xffffffffu, t, 0));
    }
    if (item >= nitems) break;
    const int64_t q = item * 32 + lane;
    const bool live = q < g.nstreams;
    int64_t c = 0, j = 0, n0 = 0, n1 = 0;
    if (live) {
        if (warm_pass) {
            const int64_t sm1 = g.S - 1;
            c = q / sm1;
            j = q - c * sm1 + 1;
            n1 = j * g.Lseg;
            n0 = max(n1 - g.warm, static_cast<int64_t>(0));
        } else {
            c = q / g.S;
            j = q - c * g.S;
            n0 = j * g.Lseg;
            n1 = min(g.T, n0 + g.Lseg);
        }
    }
    const bool from_true_state = live && n0 == 0;
    const bool do_tail = live && !warm_pass && (j == g.S - 1) && g.state_x != nullptr;
    const int tail = do_tail ? static_cast<int>(min(static_cast<int64_t>(2), n1 - n0)) : 0;
    const int64_t len = n1 - n0 - tail;  // samples filtered by the pipelined loop

    // ---- start state (DF2T) --------------------------------------------------------------
    CT s1[K], s2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s1[k] = CT(0);
        s2[k] = CT(0);
    }
    if (from_true_state) {
        if (g.state_x != nullptr) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
                const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                s1[k] = static_cast<CT>(cd.b1[k] * x1 + cd.b2[k] * x2 - cd.a1[k] * y1 - cd.a2[k] * y2);
                s2[k] = static_cast<CT>(cd.b2[k] * x1 - cd.a2[k] * y1);
            }
        }
    } else if (live && !warm_pass) {
        const CT *wsp = static_cast<const CT *>(g.ws) + (c * g.S + j);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s1[k] = wsp[(2 * k) * g.ws_stride];
            s2[k] = wsp[(2 * k + 1) * g.ws_stride];
        }
    }

    t_offx[lane] = c * g.ldx + n0;
    t_offy[lane] = c * g.ldy + n0;
    t_len[lane] = live ? len : 0;
    int64_t maxlen = live ? len : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const int64_t nch = (maxlen + CH - 1) / CH;
    __syncwarp();

    // cooperative row copies: kNvec lanes cover one row (16 bytes each), 32/kNvec rows per instruction
    constexpr int RPI = 32 / kNvec > 0 ? 32 / kNvec : 1;   // rows per instruction
    constexpr int IPR = kNvec / 32 > 0 ? kNvec / 32 : 1;   // instructions per row (rows longer than 512 B)
    const int piece = lane % kNvec;
    const int half = lane / kNvec;

    auto issue_load = [&](int64_t i, int stage) {
        unsigned char *buf = wsm + stage * (32 * kPitch);
        const int64_t base = i * CH;
        // A row is "full" (a whole chunk left), "empty" (its stream has ended) or partial.
        // All full: unconditional, fully unrolled copies (the steady state).  Full or empty:
        // the same copies under a per-row predicate.  Only a partial row -- at most one chunk
        // per stream -- needs the element-wise path.
        const unsigned fullmask = __ballot_sync(0xffffffffu, len - base >= CH);
        if (fullmask == 0xffffffffu && g.vec_ok) {
#pragma unroll
            for (int t = 0; t < 32 / RPI; ++t) {
                const int r = RPI * t + half;
#pragma unroll
                for (int u = 0; u < IPR; ++u)
                    cp_async<16>(buf + r * kPitch + (piece + 32 * u) * 16, xg + t_offx[r] + base + (piece + 32 * u) * VEC);
            }
        } else if ((fullmask | __ballot_sync(0xffffffffu, len - base <= 0)) == 0xffffffffu && g.vec_ok) {
#pragma unroll 1
            for (int t = 0; t < 32 / RPI; ++t) {
                const int r = RPI * t + half;
                if ((fullmask >> r) & 1u) {
#pragma unroll
                    for (int u = 0; u < IPR; ++u)
                        cp_async<16>(buf + r * kPitch + (piece + 32 * u) * 16, xg + t_offx[r] + base + (piece + 32 * u) * VEC);
                }
            }
        } else {
            for (int r = 0; r < 32; ++r) {
                const int64_t rem = min(t_len[r] - base, static_cast<int64_t>(CH));
                for (int e = lane; e < rem; e += 32)
                    cp_async<sizeof(IO)>(buf + r * kPitch + e * sizeof(IO), xg + t_offx[r] + base + e);
            }
        }
    };

    // ---- prologue: fill the pipeline ----------------------------------------------------
#pragma unroll
    for (int st = 0; st < kStages; ++st) {
        if (st < nch) issue_load(st, st);
        cp_async_commit();
    }

    int stage = 0;
    for (int64_t i = 0; i < nch; ++i) {
        cp_async_wait<kStages - 1>();
        __syncwarp();
        unsigned char *buf = wsm + stage * (32 * kPitch);
        const int64_t base = i * CH;
        const unsigned fullmask = __ballot_sync(0xffffffffu, len - base >= CH);
        const bool clean = fullmask == 0xffffffffu || (fullmask | __ballot_sync(0xffffffffu, len - base <= 0)) == 0xffffffffu;

        // ---- filter my row in place -------------------------------------------------------
        if (len - base >= CH) {
            Vec *row = reinterpret_cast<Vec *>(buf + lane * kPitch);
#pragma unroll UV
            for (int v = 0; v < NV; ++v) {
                Vec a = row[v];
                filter_vec<CT, K>(cf, s1, s2, a);
                row[v] = a;
            }
        } else {
            IO *row = reinterpret_cast<IO *>(buf + lane * kPitch);
            const int cnt = static_cast<int>(max(static_cast<int64_t>(0), min(len - base, static_cast<int64_t>(CH))));
            for (int e = 0; e < cnt; ++e)
                row[e] = static_cast<IO>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(row[e])));
        }
        __syncwarp();

        // ---- write the 32 rows back, coalesced --------------------------------------------
        if (!warm_pass) {
            if (fullmask == 0xffffffffu && g.vec_ok) {
#pragma unroll
                for (int t = 0; t < 32 / RPI; ++t) {
                    const int r = RPI * t + half;
#pragma unroll
                    for (int u = 0; u < IPR; ++u) {
                        const Vec v = *reinterpret_cast<const Vec *>(buf + r * kPitch + (piece + 32 * u) * 16);
                        st_stream16(yg + t_offy[r] + base + (piece + 32 * u) * VEC, v);
                    }
                }
            } else if (clean && g.vec_ok) {
#pragma unroll 1
                for (int t = 0; t < 32 / RPI; ++t) {
                    const int r = RPI * t + half;
                    if ((fullmask >> r) & 1u) {
#pragma unroll
                        for (int u = 0; u < IPR; ++u) {
                            const Vec v = *reinterpret_cast<const Vec *>(buf + r * kPitch + (piece + 32 * u) * 16);
                            st_stream16(yg + t_offy[r] + base + (piece + 32 * u) * VEC, v);
                        }
                    }
                }
            } else {
                for (int r = 0; r < 32; ++r) {
                    const int64_t rem = min(t_len[r] - base, static_cast<int64_t>(CH));
                    const IO *row = reinterpret_cast<const IO *>(buf + r * kPitch);
                    for (int e = lane; e < rem; e += 32) yg[t_offy[r] + base + e] = row[e];
                }
            }
        }
        __syncwarp();

        // ---- refill this buffer with chunk i + kStages --------------------------------------
        if (i + kStages < nch) issue_load(i + kStages, stage);
        cp_async_commit();
        stage = (stage + 1 == kStages) ? 0 : stage + 1;
    }
    cp_async_wait<0>();

    if (live && warm_pass) {
        CT *wsp = static_cast<CT *>(g.ws) + (c * g.S + j);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            wsp[(2 * k) * g.ws_stride] = s1[k];
            wsp[(2 * k + 1) * g.ws_stride] = s2[k];
        }
    }

    // ---- last two samples of the channel + DF1 state out ----------------------------------
    if (do_tail) {
        CT hx[K][2], hy[K][2];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
            // Only consulted when fewer than two samples were filtered in this call (then
            // this stream necessarily started from the caller's state: S == 1).
            hx[k][0] = static_cast<CT>(g.state_x[o]);
            hx[k][1] = static_cast<CT>(g.state_x[o + 1]);
            hy[k][0] = static_cast<CT>(g.state_y[o]);
            hy[k][1] = static_cast<CT>(g.state_y[o + 1]);
        }
        for (int e = 0; e < tail; ++e) {
            const int64_t n = n1 - tail + e;
            CT v = static_cast<CT>(xg[c * g.ldx + n]);
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
            yg[c * g.ldy + n] = static_cast<IO>(v);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
            g.state_x[o] = static_cast<double>(hx[k][0]);
            g.state_x[o + 1] = static_cast<double>(hx[k][1]);
            g.state_y[o] = static_cast<double>(hy[k][0]);
            g.state_y[o + 1] = static_cast<double>(hy[k][1]);
        }
    }
    if (g.counter == nullptr) break;
    __syncwarp();  // the per-warp tables are rewritten by the next item
  }
}

template <typename IO, typename CT, int K>
int launch_k(const SosSection *sec, Geom g, const Segmentation &seg, cudaStream_t stream) {
    SosCoef<CT, K> cf;
    SosCoefD<K> cd;
    for (int k = 0; k < K; ++k) {
        cf.b0[k] = static_cast<CT>(sec[k].b0);
        cf.b1[k] = static_cast<CT>(sec[k].b1);
        cf.b2[k] = static_cast<CT>(sec[k].b2);
        cf.na1[k] = static_cast<CT>(-sec[k].a1);
        cf.na2[k] = static_cast<CT>(-sec[k].a2);
        cd.b0[k] = sec[k].b0;
        cd.b1[k] = sec[k].b1;
        cd.b2[k] = sec[k].b2;
        cd.a1[k] = sec[k].a1;
        cd.a2[k] = sec[k].a2;
    }
    auto kern = sos_stream_kernel<IO, CT, K>;
    TFX_ENSURE_SMEM(kern, kCtaSmem);
    const int64_t per_cta = kWarps * 32;
    if (seg.S > 1) {
        Geom gw = g;
        gw.warm = seg.warm;
        gw.nstreams = g.C * (seg.S - 1);
        const int64_t grid = (gw.nstreams + per_cta - 1) / per_cta;
        kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, gw);
        TFX_CHECK_LAUNCH("sos_stream_kernel(warm-up)");
    }
    g.warm = 0;
    g.nstreams = g.C * seg.S;
    int64_t grid = (g.nstreams + per_cta - 1) / per_cta;
    const int64_t resident = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    if (seg.S > 1 && grid > resident) {
        // more items than one wave holds: persistent warps + work counter (first 8 bytes of ws)
        g.counter = static_cast<unsigned long long *>(g.ws_base);
        TFX_CUDA_TRY(cudaMemsetAsync(g.counter, 0, sizeof(unsigned long long), stream));
        grid = resident;
    }
    kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cf, cd, g);
    TFX_CHECK_LAUNCH("sos_stream_kernel");
    return TFX_OK;
}

template <typename IO, typename CT>
int launch_pass(const SosSection *sec, int k, const Geom &g, const Segmentation &seg, cudaStream_t stream) {
    switch (k) {
        case 1: return launch_k<IO, CT, 1>(sec, g, seg, stream);
        case 2: return launch_k<IO, CT, 2>(sec, g, seg, stream);
        case 3: return launch_k<IO, CT, 3>(sec, g, seg, stream);
        case 4: return launch_k<IO, CT, 4>(sec, g, seg, stream);
        case 5: return launch_k<IO, CT, 5>(sec, g, seg, stream);
        case 6: return launch_k<IO, CT, 6>(sec, g, seg, stream);
        case 7: return launch_k<IO, CT, 7>(sec, g, seg, stream);
        case 8: return launch_k<IO, CT, 8>(sec, g, seg, stream);
        default: set_error("internal: pass with %d sections", k); return TFX_EINVAL;
    }
}

int64_t stream_capacity() { return static_cast<int64_t>(sm_count()) * kWarpsPerSm * 32; }

template <typename IO>
int sos_cascade_device(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                       const double *sos_host, int K, double *state_x, double *state_y, uint32_t flags,
                       void *workspace, size_t workspace_bytes, void *stream_v) {
    TFX_REQUIRE(C >= 0 && T >= 0, "sos cascade: negative shape C=%lld T=%lld", (long long)C, (long long)T);
    TFX_REQUIRE((state_x == nullptr) == (state_y == nullptr), "sos cascade: state_x and state_y must both be given or both NULL");
    auto plan = get_sos_plan(sos_host, K);
    if (!plan) return TFX_EINVAL;
    if (C == 0 || T == 0) return TFX_OK;  // nothing to filter; state unchanged (cpu/iir_cpu.cpp loops are empty)
    TFX_REQUIRE(x != nullptr && y != nullptr, "sos cascade: NULL signal pointer");
    TFX_REQUIRE(ldx >= T && ldy >= T, "sos cascade: row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;

    uint32_t prec = flags & TFX_PREC_MASK;
    if (prec == TFX_PREC_AUTO) prec = static_cast<uint32_t>(plan->auto_prec);
    if (sizeof(IO) == 8) prec = TFX_PREC_F64;
    TFX_REQUIRE(prec == TFX_PREC_F32 || prec == TFX_PREC_F64, "sos cascade: bad precision flag");
    const bool no_split = (flags & TFX_NO_SPLIT) != 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

    for (size_t pi = 0; pi < plan->passes.size(); ++pi) {
        const SosPass &p = plan->passes[pi];
        const int64_t warm_needed = prec == TFX_PREC_F32 ? p.warm_f32 : (sizeof(IO) == 4 ? p.warm_f64_io32 : p.warm_f64_io64);
        const IO *px = pi == 0 ? x : y;
        const int64_t pldx = pi == 0 ? ldx : ldy;
        // TFX_PREC_AUTO picked float64 but only some sections need it: mixed-precision tile kernel
        if constexpr (sizeof(IO) == 4) {
            const unsigned plan_mask = static_cast<unsigned>((plan->mixed_mask >> p.k0) & ((1ull << p.k) - 1ull));
            const unsigned local_mask = plan_mask == 0u ? 0u : tile_mixed_cover(p.k, plan_mask);  // widened to an instantiated mask
            const bool mixed = (flags & TFX_PREC_MASK) == TFX_PREC_AUTO && plan->auto_prec == TFX_PREC_F64 && !(flags & TFX_NO_TILE) &&
                               tile_path_ok(C) && (plan_mask == 0u || local_mask != 0u);
            if (mixed) {
                const int64_t lanes = (C + 31) / 32 * 32;
                const Segmentation seg = choose_segmentation(lanes, T, p.warm_f64_io32, tile_stream_capacity(), no_split, kOversub);
                if (seg.S > 1) {
                    const size_t need = kWsHeader + static_cast<size_t>(2 * p.k) * static_cast<size_t>(C * seg.S) * 8;
                    if (workspace == nullptr || workspace_bytes < need) {
                        set_error("sos cascade: workspace of %zu bytes needed, %zu given (query tfx_sos_cascade_workspace_bytes)", need,
                                  workspace_bytes);
                        return TFX_EWORKSPACE;
                    }
                }
                double *psx = state_x ? state_x + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
                double *psy = state_y ? state_y + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
                if (local_mask == 0u)
                    rc = launch_tile_pass<float, float>(px, y, C, T, pldx, ldy, plan->sec.data() + p.k0, p.k, seg, workspace, psx, psy, stream);
                else
                    rc = launch_tile_pass_mixed(px, y, C, T, pldx, ldy, plan->sec.data() + p.k0, p.k, local_mask, seg, workspace, psx, psy, stream);
                if (rc != TFX_OK) return rc;
                continue;
            }
        }
        if (!(flags & TFX_NO_TILE) && tile_path_ok(C)) {
            // channel-tile kernel: a warp is 32 consecutive channels x one time segment
            const int64_t lanes = (C + 31) / 32 * 32;
            const Segmentation seg = choose_segmentation(lanes, T, warm_needed, tile_stream_capacity(), no_split, kOversub);
            if (seg.S > 1) {
                const size_t need = kWsHeader + static_cast<size_t>(2 * p.k) * static_cast<size_t>(C * seg.S) * (prec == TFX_PREC_F32 ? 4 : 8);
                if (workspace == nullptr || workspace_bytes < need) {
                    set_error("sos cascade: workspace of %zu bytes needed, %zu given (query tfx_sos_cascade_workspace_bytes)", need,
                              workspace_bytes);
                    return TFX_EWORKSPACE;
                }
            }
            double *psx = state_x ? state_x + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
            double *psy = state_y ? state_y + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
            if (prec == TFX_PREC_F32) {
                if constexpr (sizeof(IO) == 4)
                    rc = launch_tile_pass<IO, float>(px, y, C, T, pldx, ldy, plan->sec.data() + p.k0, p.k, seg, workspace, psx, psy, stream);
                else
                    rc = TFX_EINVAL;
            } else {
                rc = launch_tile_pass<IO, double>(px, y, C, T, pldx, ldy, plan->sec.data() + p.k0, p.k, seg, workspace, psx, psy, stream);
            }
            if (rc != TFX_OK) return rc;
            continue;
        }
        // stream-per-lane kernel (few channels, or on request)
        const Segmentation seg = choose_segmentation(C, T, warm_needed, stream_capacity(), no_split, kOversub);
        Geom g{};
        g.x = pi == 0 ? static_cast<const void *>(x) : static_cast<const void *>(y);
        g.y = y;
        g.ldx = pi == 0 ? ldx : ldy;
        g.ldy = ldy;
        g.C = C;
        g.T = T;
        g.S = seg.S;
        g.Lseg = seg.Lseg;
        g.ws_base = workspace;
        g.ws = workspace ? static_cast<unsigned char *>(workspace) + kWsHeader : nullptr;
        g.ws_stride = C * seg.S;
        g.counter = nullptr;
        g.state_x = state_x ? state_x + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
        g.state_y = state_y ? state_y + static_cast<int64_t>(p.k0) * C * 2 : nullptr;
        const size_t esz = sizeof(IO);
        g.vec_ok = (reinterpret_cast<uintptr_t>(g.x) % 16 == 0) && (reinterpret_cast<uintptr_t>(g.y) % 16 == 0) &&
                   ((g.ldx * esz) % 16 == 0) && ((g.ldy * esz) % 16 == 0) && (seg.S == 1 || (seg.Lseg * esz) % 16 == 0);
        if (seg.S > 1) {
            const size_t need = kWsHeader + static_cast<size_t>(2 * p.k) * static_cast<size_t>(C * seg.S) * (prec == TFX_PREC_F32 ? 4 : 8);
            if (workspace == nullptr || workspace_bytes < need) {
                set_error("sos cascade: workspace of %zu bytes needed, %zu given (query tfx_sos_cascade_workspace_bytes)", need,
                          workspace_bytes);
                return TFX_EWORKSPACE;
            }
        }
        if (prec == TFX_PREC_F32) {
            if constexpr (sizeof(IO) == 4) {
                rc = launch_pass<IO, float>(plan->sec.data() + p.k0, p.k, g, seg, stream);
            } else {
                rc = TFX_EINVAL;
            }
        } else {
            rc = launch_pass<IO, double>(plan->sec.data() + p.k0, p.k, g, seg, stream);
        }
        if (rc != TFX_OK) return rc;
    }
    return TFX_OK;
}

}  // namespace
}  // namespace tfx

extern "C" {

size_t tfx_sos_cascade_workspace_bytes(int64_t C, int64_t T, int K) {
    (void)T;
    if (C <= 0 || K <= 0) return 0;
    // S > 1 only when C*S <= one wave of streams; state is 2 values per fused section.
    const int64_t streams = std::max<int64_t>(tfx::stream_capacity(), tfx::tile_stream_capacity()) * tfx::kOversub + C + 128;
    const int kf = K < TFX_SOS_MAX_FUSED ? K : TFX_SOS_MAX_FUSED;
    return tfx::kWsHeader + static_cast<size_t>(2 * kf) * static_cast<size_t>(streams) * 8 + 256;
}

int tfx_sos_cascade_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        const double *sos_host, int K, double *state_x, double *state_y, uint32_t flags,
                        void *workspace, size_t workspace_bytes, void *stream) {
    return tfx::sos_cascade_device<float>(x, y, C, T, ldx, ldy, sos_host, K, state_x, state_y, flags, workspace,
                                          workspace_bytes, stream);
}

int tfx_sos_cascade_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        const double *sos_host, int K, double *state_x, double *state_y, uint32_t flags,
                        void *workspace, size_t workspace_bytes, void *stream) {
    return tfx::sos_cascade_device<double>(x, y, C, T, ldx, ldy, sos_host, K, state_x, state_y, flags, workspace,
                                           workspace_bytes, stream);
}

uint64_t tfx_sos_mixed_mask(const double *sos_host, int K, double *mixed_rel_err) {
    auto plan = tfx::get_sos_plan(sos_host, K);
    if (!plan) return 0;
    if (mixed_rel_err) *mixed_rel_err = plan->mixed_rel_err;
    return plan->auto_prec == TFX_PREC_F64 ? plan->mixed_mask : 0;
}

void tfx_plan_segmentation(int64_t lanes, int64_t T, int64_t warm_needed, int64_t capacity, int oversub, int64_t *segments,
                           int64_t *segment_len, int64_t *warm) {
    const tfx::Segmentation g = tfx::choose_segmentation(lanes, T, warm_needed, capacity, false, oversub < 1 ? 1 : oversub);
    if (segments) *segments = g.S;
    if (segment_len) *segment_len = g.Lseg;
    if (warm) *warm = g.warm;
}

int tfx_sos_auto_precision(const double *sos_host, int K, double *probe_rel_err) {
    auto plan = tfx::get_sos_plan(sos_host, K);
    if (!plan) return TFX_EINVAL;
    if (probe_rel_err) *probe_rel_err = plan->probe_rel_err;
    return plan->auto_prec;
}

}  // extern "C"
