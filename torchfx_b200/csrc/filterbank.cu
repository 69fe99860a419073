// filterbank.cu -- N parallel SOS filters over the SAME input in one launch.
//
// Replaces the reference's Python loops: LogFilterBank.forward (filter/filterbank.py:183-185,
// n_bands native calls + torch.stack) and ParallelFilterCombination.forward
// (filter/__base.py:1019-1026, N native calls + N temporaries + N adds).
//
// Dispatch (filterbank_device): with enough channels to fill 32-lane groups and float32 I/O the bank runs on the
// lanes = channels kernels -- SUM banks of <= 8 sections on the cascade tile kernel's parallel topology
// (bank_tile.cu, bank_tile_mixed.cu), STACK banks and larger SUM banks on bank_stack_kernel (bank_stack.cu).
// Few channels, float64 I/O and SUM banks of more than 32 bands take the kernel in THIS file:
//
// Layout: BAND PER LANE.  A warp serves 32/LB streams (stream = channel x time segment,
// exactly as in sos_cascade.cu); within a stream LB = next_pow2(N) lanes each own one band:
// the band's coefficients and DF2T state live in that lane's registers.  Per 256-byte input
// chunk of a stream all its lanes read the same samples from shared memory (broadcast), so
// x costs 4/N bytes per lane-sample, and write their band's output row into a 32-row
// shared-memory tile that leaves the SM as 256-byte coalesced rows:
//   STACK: row (band b, channel c) -> y[b, c, t..t+64)             4*(1+1/N) B per lane-sample
//   SUM  : the tile is reduced over bands in band order (the reference's `+=` order)
//          and ONE row per stream is stored                        8 B per channel-sample
// Time segmentation, warm-up launch, DF1 state hand-over: identical to sos_cascade.cu.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "bank_tile.h"
#include "common.cuh"
#include "sos_kernels.h"
#include "sos_plan.h"
#include "stream_common.cuh"

namespace tfx {
namespace {

constexpr int kWarps = 4;
constexpr int kStages = 3;
constexpr int kMaxSpw = 16;  // streams per warp (LB >= 2)
constexpr int kOutBytes = 32 * kPitch;
constexpr int kMaxKb = 4;
constexpr int kMaxCtasPerSm = 5;  // register budget: __launch_bounds__(128, 5)
// Shared memory of one warp depends on how many streams it serves (32 / lanes-per-stream):
// [kStages][spw] input rows, one 32-row output tile, offset tables.
__host__ __device__ constexpr int warp_smem_bytes(int spw) { return kStages * spw * kPitch + kOutBytes + (32 + 2 * spw) * 8; }
inline int ctas_per_sm(int spw) {
    const int n = kSmemPerSm / (kWarps * warp_smem_bytes(spw) + 1024);
    return n < 1 ? 1 : (n > kMaxCtasPerSm ? kMaxCtasPerSm : n);
}

template <int KB>
struct BankCoef {
    double b0[32][KB], b1[32][KB], b2[32][KB], a1[32][KB], a2[32][KB];
};

struct BankGeom {
    const void *x;
    void *y;
    int64_t ldx, ldy, ldb, C, T;
    int64_t S, Lseg, warm, nstreams;
    void *ws;
    int64_t ws_stride;
    double *state_x;  // [N(group), KB, C, 2] of THIS band group
    double *state_y;
    int n_bands;      // bands in this launch (<= 32)
    int band_id[32];  // global band index of local band b (y row block and state block)
    int lb_shift;     // log2(lanes per stream)
    int mode;         // TFX_BANK_STACK / TFX_BANK_SUM
    int accumulate;   // SUM: add onto what y already holds (band groups after the first)
    int vec_ok;
};

template <typename IO, typename CT, int KB>
__global__ void __launch_bounds__(kWarps * 32, sizeof(CT) == 8 ? 3 : kMaxCtasPerSm)
bank_stream_kernel(const __grid_constant__ BankCoef<KB> cd, const __grid_constant__ BankGeom g) {
    using Tr = IoTraits<IO>;
    using Vec = typename Tr::Vec;
    constexpr int CH = Tr::CHUNK;
    constexpr int VEC = Tr::VEC;
    constexpr int NV = kNvec;
    static_assert(kNvec <= 32, "filterbank assumes rows of at most 512 bytes");
    constexpr int RPI = 32 / kNvec;  // rows per cooperative instruction

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int LB = 1 << g.lb_shift;
    const int SPW = 32 >> g.lb_shift;
    const int in_stage = SPW * kPitch;
    unsigned char *wsm = smem_raw + warp * warp_smem_bytes(SPW);
    unsigned char *in_bufs = wsm;
    unsigned char *out_tile = wsm + kStages * in_stage;
    int64_t *t_offy = reinterpret_cast<int64_t *>(out_tile + kOutBytes);  // per lane-row
    int64_t *t_offx = t_offy + 32;                                        // per stream slot
    int64_t *t_len = t_offx + SPW;

    const IO *__restrict__ xg = static_cast<const IO *>(g.x);
    IO *__restrict__ yg = static_cast<IO *>(g.y);

    const int slot = lane >> g.lb_shift;
    const int band = lane & (LB - 1);
    const bool band_ok = band < g.n_bands;
    const int64_t gband = g.band_id[band_ok ? band : 0];

    const int64_t q = (static_cast<int64_t>(blockIdx.x) * kWarps + warp) * SPW + slot;
    const bool live = q < g.nstreams;
    const bool warm_pass = g.warm > 0;
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
    const int64_t len = live ? n1 - n0 - tail : 0;

    // ---- my band's coefficients and start state -------------------------------------------
    CT b0[KB], b1[KB], b2[KB], na1[KB], na2[KB], s1[KB], s2[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        b0[k] = static_cast<CT>(cd.b0[band][k]);
        b1[k] = static_cast<CT>(cd.b1[band][k]);
        b2[k] = static_cast<CT>(cd.b2[band][k]);
        na1[k] = static_cast<CT>(-cd.a1[band][k]);
        na2[k] = static_cast<CT>(-cd.a2[band][k]);
        s1[k] = CT(0);
        s2[k] = CT(0);
    }
    if (band_ok && from_true_state) {
        if (g.state_x != nullptr) {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int64_t o = ((gband * KB + k) * g.C + c) * 2;
                const double x1 = g.state_x[o], x2 = g.state_x[o + 1];
                const double y1 = g.state_y[o], y2 = g.state_y[o + 1];
                s1[k] = static_cast<CT>(cd.b1[band][k] * x1 + cd.b2[band][k] * x2 - cd.a1[band][k] * y1 - cd.a2[band][k] * y2);
                s2[k] = static_cast<CT>(cd.b2[band][k] * x1 - cd.a2[band][k] * y1);
            }
        }
    } else if (band_ok && live && !warm_pass) {
        const CT *wsp = static_cast<const CT *>(g.ws) + (c * g.S + j);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            s1[k] = wsp[((band * KB + k) * 2) * g.ws_stride];
            s2[k] = wsp[((band * KB + k) * 2 + 1) * g.ws_stride];
        }
    }

    auto step = [&](CT v) -> CT {
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const CT y = fma_rn(b0[k], v, s1[k]);
            s1[k] = fma_rn(na1[k], y, fma_rn(b1[k], v, s2[k]));
            s2[k] = fma_rn(na2[k], y, b2[k] * v);
            v = y;
        }
        return v;
    };

    if (band == 0) {
        t_offx[slot] = c * g.ldx + n0;
        t_len[slot] = len;
    }
    t_offy[lane] = gband * g.ldb + c * g.ldy + n0;
    int64_t maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const int64_t nch = (maxlen + CH - 1) / CH;
    __syncwarp();

    auto issue_load = [&](int64_t i, int stage) {
        unsigned char *buf = in_bufs + stage * in_stage;
        const int64_t base = i * CH;
        // steady state: every stream of the warp has a whole chunk left -> unconditional copies
        if (g.vec_ok && __all_sync(0xffffffffu, len - base >= CH || len - base <= 0)) {
            for (int idx = lane; idx < SPW * kNvec; idx += 32) {
                const int r = idx / kNvec, piece = idx % kNvec;
                if (t_len[r] > base) cp_async<16>(buf + r * kPitch + piece * 16, xg + t_offx[r] + base + piece * VEC);
            }
            return;
        }
        for (int idx = lane; idx < SPW * kNvec; idx += 32) {
            const int r = idx / kNvec, piece = idx % kNvec;
            const int64_t rem = t_len[r] - base;
            const IO *src = xg + t_offx[r] + base + piece * VEC;
            unsigned char *dst = buf + r * kPitch + piece * 16;
            if (g.vec_ok && rem >= (piece + 1) * VEC) {
                cp_async<16>(dst, src);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (piece * VEC + v < rem) cp_async<sizeof(IO)>(dst + v * sizeof(IO), src + v);
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
        const unsigned char *buf = in_bufs + stage * in_stage;
        const int64_t base = i * CH;
        const int cnt = static_cast<int>(max(static_cast<int64_t>(0), min(len - base, static_cast<int64_t>(CH))));

        // ---- every lane filters its band over the stream's chunk --------------------------
        if (band_ok) {
            if (cnt == CH) {
                const Vec *xin = reinterpret_cast<const Vec *>(buf + slot * kPitch);
                Vec *row = reinterpret_cast<Vec *>(out_tile + lane * kPitch);
#pragma unroll 4
                for (int v = 0; v < NV; ++v) {
                    Vec a = xin[v];
                    if constexpr (VEC == 4) {
                        a.x = static_cast<IO>(step(static_cast<CT>(a.x)));
                        a.y = static_cast<IO>(step(static_cast<CT>(a.y)));
                        a.z = static_cast<IO>(step(static_cast<CT>(a.z)));
                        a.w = static_cast<IO>(step(static_cast<CT>(a.w)));
                    } else {
                        a.x = static_cast<IO>(step(static_cast<CT>(a.x)));
                        a.y = static_cast<IO>(step(static_cast<CT>(a.y)));
                    }
                    row[v] = a;
                }
            } else {
                const IO *xin = reinterpret_cast<const IO *>(buf + slot * kPitch);
                IO *row = reinterpret_cast<IO *>(out_tile + lane * kPitch);
                for (int e = 0; e < cnt; ++e) row[e] = static_cast<IO>(step(static_cast<CT>(xin[e])));
            }
        }
        __syncwarp();

        // ---- leave the SM as coalesced rows -----------------------------------------------
        if (!warm_pass) {
            const bool all_full = g.vec_ok && __all_sync(0xffffffffu, len - base >= CH || len - base <= 0);
            if (g.mode == TFX_BANK_STACK && all_full) {
                // steady state: 16 unconditional coalesced row stores (rows of absent bands / streams skipped)
                const int piece = lane % kNvec, half = lane / kNvec;
#pragma unroll
                for (int t = 0; t < 32 / RPI; ++t) {
                    const int r = RPI * t + half;
                    if ((r & (LB - 1)) < g.n_bands && t_len[r >> g.lb_shift] > base)
                        st_stream16(yg + t_offy[r] + base + piece * VEC, *reinterpret_cast<const Vec *>(out_tile + r * kPitch + piece * 16));
                }
            } else if (g.mode == TFX_BANK_STACK) {
                const int piece = lane % kNvec, half = lane / kNvec;
#pragma unroll 1
                for (int t = 0; t < 32 / RPI; ++t) {
                    const int r = RPI * t + half;
                    if ((r & (LB - 1)) >= g.n_bands) continue;
                    const int64_t rem = t_len[r >> g.lb_shift] - base;
                    const unsigned char *src = out_tile + r * kPitch + piece * 16;
                    IO *dst = yg + t_offy[r] + base + piece * VEC;
                    if (g.vec_ok && rem >= (piece + 1) * VEC) {
                        st_stream16(dst, *reinterpret_cast<const Vec *>(src));
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            if (piece * VEC + v < rem) dst[v] = reinterpret_cast<const IO *>(src)[v];
                    }
                }
            } else {
                for (int s = 0; s < SPW; ++s) {
                    const int64_t rem = min(t_len[s] - base, static_cast<int64_t>(CH));
                    IO *dst = yg + (t_offy[s << g.lb_shift]) + base;  // band-0 row of the slot: ldb term is 0
                    for (int e = lane; e < rem; e += 32) {
                        IO acc = g.accumulate ? dst[e] : IO(0);
                        for (int b = 0; b < g.n_bands; ++b)
                            acc += reinterpret_cast<const IO *>(out_tile + ((s << g.lb_shift) + b) * kPitch)[e];
                        dst[e] = acc;
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

    if (warm_pass) {
        if (live && band_ok) {
            CT *wsp = static_cast<CT *>(g.ws) + (c * g.S + j);
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                wsp[((band * KB + k) * 2) * g.ws_stride] = s1[k];
                wsp[((band * KB + k) * 2 + 1) * g.ws_stride] = s2[k];
            }
        }
        return;
    }

    // ---- last two samples of each channel + DF1 state out (warp-uniform control flow) -----
    if (__any_sync(0xffffffffu, do_tail)) {
        const bool mine = do_tail && band_ok;
        CT hx[KB][2], hy[KB][2];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            hx[k][0] = hx[k][1] = hy[k][0] = hy[k][1] = CT(0);
            if (mine) {
                const int64_t o = ((gband * KB + k) * g.C + c) * 2;
                hx[k][0] = static_cast<CT>(g.state_x[o]);
                hx[k][1] = static_cast<CT>(g.state_x[o + 1]);
                hy[k][0] = static_cast<CT>(g.state_y[o]);
                hy[k][1] = static_cast<CT>(g.state_y[o + 1]);
            }
        }
        for (int e = 0; e < 2; ++e) {
            const bool act = mine && e < tail;
            const int64_t n = n1 - tail + e;
            CT v = act ? static_cast<CT>(xg[c * g.ldx + n]) : CT(0);
            if (act) {
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    const CT y = fma_rn(b0[k], v, s1[k]);
                    s1[k] = fma_rn(na1[k], y, fma_rn(b1[k], v, s2[k]));
                    s2[k] = fma_rn(na2[k], y, b2[k] * v);
                    hx[k][1] = hx[k][0];
                    hx[k][0] = v;
                    hy[k][1] = hy[k][0];
                    hy[k][0] = y;
                    v = y;
                }
            }
            if (g.mode == TFX_BANK_STACK) {
                if (act) yg[gband * g.ldb + c * g.ldy + n] = static_cast<IO>(v);
            } else {
                // band-ordered sum, identical to the tile reduction above
                IO acc = IO(0);
                const IO mine_v = act ? static_cast<IO>(v) : IO(0);
                for (int b = 0; b < g.n_bands; ++b) acc += __shfl_sync(0xffffffffu, mine_v, (slot << g.lb_shift) + b);
                if (act && band == 0) {
                    IO *dst = yg + c * g.ldy + n;
                    *dst = g.accumulate ? *dst + acc : acc;
                }
            }
        }
        if (mine) {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int64_t o = ((gband * KB + k) * g.C + c) * 2;
                g.state_x[o] = static_cast<double>(hx[k][0]);
                g.state_x[o + 1] = static_cast<double>(hx[k][1]);
                g.state_y[o] = static_cast<double>(hy[k][0]);
                g.state_y[o + 1] = static_cast<double>(hy[k][1]);
            }
        }
    }
}

template <typename IO, typename CT, int KB>
int launch_bank(const BankCoef<KB> &cd, BankGeom g, const Segmentation &seg, cudaStream_t stream) {
    auto kern = bank_stream_kernel<IO, CT, KB>;
    TFX_ENSURE_SMEM(kern, kWarps * warp_smem_bytes(kMaxSpw));
    const int kCtaSmem = kWarps * warp_smem_bytes(32 >> g.lb_shift);
    const int64_t per_cta = static_cast<int64_t>(kWarps) * (32 >> g.lb_shift);
    if (seg.S > 1) {
        BankGeom gw = g;
        gw.warm = seg.warm;
        gw.nstreams = g.C * (seg.S - 1);
        const int64_t grid = (gw.nstreams + per_cta - 1) / per_cta;
        kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cd, gw);
        TFX_CHECK_LAUNCH("bank_stream_kernel(warm-up)");
    }
    g.warm = 0;
    g.nstreams = g.C * seg.S;
    const int64_t grid = (g.nstreams + per_cta - 1) / per_cta;
    kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(cd, g);
    TFX_CHECK_LAUNCH("bank_stream_kernel");
    return TFX_OK;
}

template <typename IO, typename CT, int KB>
int run_group(const std::vector<std::shared_ptr<const SosPlan>> &plans, const int *bands, int nb, BankGeom g,
              const Segmentation &seg, cudaStream_t stream) {
    BankCoef<KB> cd;
    for (int b = 0; b < 32; ++b)
        for (int k = 0; k < KB; ++k) {
            const bool have = b < nb;
            const SosSection id{1.0, 0.0, 0.0, 0.0, 0.0};
            const SosSection &s = have ? plans[bands[b]]->sec[k] : id;
            cd.b0[b][k] = s.b0;
            cd.b1[b][k] = s.b1;
            cd.b2[b][k] = s.b2;
            cd.a1[b][k] = s.a1;
            cd.a2[b][k] = s.a2;
        }
    return launch_bank<IO, CT, KB>(cd, g, seg, stream);
}

template <typename IO, typename CT>
int run_group_kb(int KB, const std::vector<std::shared_ptr<const SosPlan>> &plans, const int *bands, int nb, const BankGeom &g,
                 const Segmentation &seg, cudaStream_t stream) {
    switch (KB) {
        case 1: return run_group<IO, CT, 1>(plans, bands, nb, g, seg, stream);
        case 2: return run_group<IO, CT, 2>(plans, bands, nb, g, seg, stream);
        case 3: return run_group<IO, CT, 3>(plans, bands, nb, g, seg, stream);
        case 4: return run_group<IO, CT, 4>(plans, bands, nb, g, seg, stream);
        default: set_error("filterbank: Kb must be in [1, %d]", kMaxKb); return TFX_EINVAL;
    }
}

int lanes_shift(int nb) {
    int sh = 1;  // at least 2 lanes per stream
    while ((1 << sh) < nb) ++sh;
    return sh;
}

template <typename IO>
int filterbank_device(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t ldb,
                      const double *sos_host, int N, int Kb, int mode, double *state_x, double *state_y, uint32_t flags,
                      void *workspace, size_t workspace_bytes, void *stream_v) {
    TFX_REQUIRE(C >= 0 && T >= 0, "filterbank: negative shape");
    TFX_REQUIRE(N >= 1 && Kb >= 1 && Kb <= kMaxKb, "filterbank: need N >= 1 and 1 <= Kb <= %d (got N=%d Kb=%d)", kMaxKb, N, Kb);
    TFX_REQUIRE(mode == TFX_BANK_STACK || mode == TFX_BANK_SUM, "filterbank: bad mode %d", mode);
    TFX_REQUIRE((state_x == nullptr) == (state_y == nullptr), "filterbank: state_x and state_y must both be given or both NULL");
    TFX_REQUIRE(sos_host != nullptr, "filterbank: NULL coefficients");
    std::vector<std::shared_ptr<const SosPlan>> plans(N);
    for (int b = 0; b < N; ++b) {
        plans[b] = get_sos_plan(sos_host + static_cast<size_t>(b) * Kb * 6, Kb);
        if (!plans[b]) return TFX_EINVAL;
    }
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x != nullptr && y != nullptr && static_cast<const void *>(x) != static_cast<const void *>(y),
                "filterbank: NULL or aliased signal pointers (not in place)");
    TFX_REQUIRE(ldx >= T && ldy >= T, "filterbank: row stride smaller than T");
    TFX_REQUIRE(mode == TFX_BANK_SUM || ldb >= (C - 1) * ldy + T, "filterbank: band stride ldb too small for [N, C, T]");
    int rc = require_device();
    if (rc != TFX_OK) return rc;

    // SUM banks of up to 8 sections with enough channels: the channel-tile kernel with the parallel
    // topology (bank_tile.cu) -- x read once, y written once, branch states in registers.
    if constexpr (sizeof(IO) == 4) {
        if (mode == TFX_BANK_SUM && !(flags & TFX_NO_TILE) && bank_sum_tile_ok(N, Kb, C)) {
            uint32_t prec = flags & TFX_PREC_MASK;
            int n64 = 0;  // AUTO: branches [0, n64) need float64 and the others do not -> mixed kernel
            if (prec == TFX_PREC_AUTO) {
                prec = TFX_PREC_F32;
                bool prefix = true;
                for (int b = 0; b < N; ++b)
                    if (plans[b]->auto_prec == TFX_PREC_F64) {
                        prec = TFX_PREC_F64;
                        prefix = prefix && b == n64;
                        ++n64;
                    }
                if (!prefix || !bank_sum_mixed_ok(N, Kb, n64)) n64 = 0;
            }
            TFX_REQUIRE(prec == TFX_PREC_F32 || prec == TFX_PREC_F64, "filterbank: bad precision flag");
            int64_t warm_needed = 0;
            std::vector<SosSection> sec;
            for (int b = 0; b < N; ++b) {
                const SosPass &p = plans[b]->passes[0];
                const int64_t w = prec == TFX_PREC_F32 ? p.warm_f32 : p.warm_f64_io32;
                warm_needed = (w < 0 || warm_needed < 0) ? -1 : std::max(warm_needed, w);
                for (int k = 0; k < Kb; ++k) sec.push_back(plans[b]->sec[k]);
            }
            const int64_t lanes = (C + 31) / 32 * 32;
            const Segmentation seg =
                choose_segmentation(lanes, T, warm_needed, bank_tile_stream_capacity(), (flags & TFX_NO_SPLIT) != 0, kOversub);
            if (seg.S > 1) {
                const size_t need = kWsHeader + static_cast<size_t>(2 * N * Kb) * static_cast<size_t>(C * seg.S) * (prec == TFX_PREC_F32 ? 4 : 8);
                if (workspace == nullptr || workspace_bytes < need) {
                    set_error("filterbank: workspace of %zu bytes needed, %zu given (query tfx_filterbank_workspace_bytes)", need,
                              workspace_bytes);
                    return TFX_EWORKSPACE;
                }
            }
            cudaStream_t st = static_cast<cudaStream_t>(stream_v);
            if (n64 > 0)
                return launch_bank_sum_tile_mixed(x, y, C, T, ldx, ldy, sec.data(), N, Kb, n64, seg, workspace, state_x, state_y, st);
            if (prec == TFX_PREC_F32)
                return launch_bank_sum_tile<float>(x, y, C, T, ldx, ldy, sec.data(), N, Kb, seg, workspace, state_x, state_y, st);
            return launch_bank_sum_tile<double>(x, y, C, T, ldx, ldy, sec.data(), N, Kb, seg, workspace, state_x, state_y, st);
        }
    }

    // STACK banks with enough channels: lanes = channels, per-band precision and warm-up (bank_stack.cu)
    if constexpr (sizeof(IO) == 4) {
        // (also SUM banks too large for the register-resident parallel topology above, up to 32 bands)
        bool stack_kernel = (mode == TFX_BANK_STACK || N <= 32) && !(flags & TFX_NO_TILE) && bank_stack_tile_ok(N, Kb, C);
        if (mode == TFX_BANK_SUM && (flags & TFX_BANK_STRICT_ORDER)) stack_kernel = false;  // child-order sum: band-per-lane kernel
        if (stack_kernel && !(flags & TFX_FORCE_TILE)) {  // enough (channel group x segment x band split) items to fill the GPU?
            int64_t warm_max = 0;
            for (int b = 0; b < std::min(N, 32); ++b) {
                const int64_t w = plans[b]->passes[0].warm_f32;
                warm_max = (w < 0 || warm_max < 0) ? -1 : std::max(warm_max, w);
            }
            stack_kernel = bank_stack_worthwhile(C, T, std::min(N, 32), Kb, warm_max, mode == TFX_BANK_SUM, (flags & TFX_NO_SPLIT) != 0);
        }
        if (stack_kernel) {
            const uint32_t want = flags & TFX_PREC_MASK;
            TFX_REQUIRE(want == TFX_PREC_AUTO || want == TFX_PREC_F32 || want == TFX_PREC_F64, "filterbank: bad precision flag");
            for (int lo = 0; lo < N; lo += 32) {
                const int nb = std::min(32, N - lo);
                std::vector<SosSection> sec;
                int band_id[32];
                int64_t warm_b[32];
                uint32_t mask = 0;
                for (int b = 0; b < nb; ++b) {
                    const SosPlan &pl = *plans[lo + b];
                    const uint32_t pb = want == TFX_PREC_AUTO ? static_cast<uint32_t>(pl.auto_prec) : want;
                    if (pb == TFX_PREC_F64) mask |= 1u << b;
                    // float32 I/O: a start state exact to 2^-30 is below the output rounding whatever the recurrence precision
                    (void)pb;
                    warm_b[b] = pl.passes[0].warm_f32;
                    band_id[b] = lo + b;
                    for (int k = 0; k < Kb; ++k) sec.push_back(pl.sec[k]);
                }
                rc = launch_bank_stack(x, y, C, T, ldx, ldy, ldb, sec.data(), band_id, warm_b, mask, nb, Kb, mode == TFX_BANK_SUM,
                                       (flags & TFX_NO_SPLIT) != 0, workspace, workspace_bytes, state_x, state_y,
                                       static_cast<cudaStream_t>(stream_v));
                if (rc != TFX_OK) return rc;
            }
            return TFX_OK;
        }
    }

    // Precision per band.  With TFX_PREC_AUTO the bank is split by the per-band probe: bands whose
    // float32 recurrence is accurate run in a float32 launch, the others (low corners) in a
    // float64 launch -- the lanes of one launch execute one instruction stream, so mixing would
    // make every band pay for float64.  (In SUM mode the f32 group is accumulated first; the
    // summation order therefore differs from the reference's child order by rounding only.)
    const uint32_t want = flags & TFX_PREC_MASK;
    const bool no_split = (flags & TFX_NO_SPLIT) != 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const size_t esz = sizeof(IO);
    std::vector<int> list32, list64;
    for (int b = 0; b < N; ++b) {
        uint32_t pb = want == TFX_PREC_AUTO ? static_cast<uint32_t>(plans[b]->auto_prec) : want;
        if (esz == 8) pb = TFX_PREC_F64;
        (pb == TFX_PREC_F32 ? list32 : list64).push_back(b);
    }
    // Splitting only pays when the float64 launch gets narrower (fewer lanes per stream) than the
    // whole bank would be: otherwise one float64 launch reads x once and fills the lanes better
    // (measured: 32 bands, 19 of them f64: 30.1 ms unsplit vs 34.7 ms split).
    // SUM mode never splits: a second launch would re-read x and read-modify-write y (measured 52 ms
    // split vs 21 ms unsplit for 8 bands over 1024 channels).
    if (!list32.empty() && !list64.empty() &&
        (mode == TFX_BANK_SUM || (N <= 32 && lanes_shift(static_cast<int>(list64.size())) >= lanes_shift(N)))) {
        list64.clear();
        list32.clear();
        for (int b = 0; b < N; ++b) list64.push_back(b);
    }
    bool first_launch = true;
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int> &list = pass == 0 ? list32 : list64;
        const uint32_t prec = pass == 0 ? TFX_PREC_F32 : TFX_PREC_F64;
        for (size_t lo = 0; lo < list.size(); lo += 32) {
            const int nb = static_cast<int>(std::min<size_t>(32, list.size() - lo));
            const int *bands = list.data() + lo;
            int64_t warm_needed = 0;
            for (int b = 0; b < nb; ++b) {
                const SosPass &p = plans[bands[b]]->passes[0];
                const int64_t w = prec == TFX_PREC_F32 ? p.warm_f32 : (esz == 4 ? p.warm_f64_io32 : p.warm_f64_io64);
                if (w < 0) {
                    warm_needed = -1;
                    break;
                }
                warm_needed = std::max(warm_needed, w);
            }
            const int sh = lanes_shift(nb);
            const int spw = 32 >> sh;
            const int64_t capacity = static_cast<int64_t>(sm_count()) * ctas_per_sm(spw) * kWarps * spw;
            const Segmentation seg = choose_segmentation(C, T, warm_needed, capacity, no_split);
            BankGeom g{};
            g.x = x;
            g.y = y;
            g.ldx = ldx;
            g.ldy = ldy;
            g.ldb = mode == TFX_BANK_STACK ? ldb : 0;
            g.C = C;
            g.T = T;
            g.S = seg.S;
            g.Lseg = seg.Lseg;
            g.ws = workspace;
            g.ws_stride = C * seg.S;
            g.state_x = state_x;
            g.state_y = state_y;
            g.n_bands = nb;
            for (int b = 0; b < 32; ++b) g.band_id[b] = bands[b < nb ? b : 0];
            g.lb_shift = sh;
            g.mode = mode;
            g.accumulate = (mode == TFX_BANK_SUM && !first_launch) ? 1 : 0;
            g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
                       ((ldx * esz) % 16 == 0) && ((ldy * esz) % 16 == 0) && ((g.ldb * esz) % 16 == 0) &&
                       (seg.S == 1 || (seg.Lseg * esz) % 16 == 0);
            if (seg.S > 1) {
                const size_t need = static_cast<size_t>(2 * Kb * nb) * static_cast<size_t>(C * seg.S) * (prec == TFX_PREC_F32 ? 4 : 8);
                if (workspace == nullptr || workspace_bytes < need) {
                    set_error("filterbank: workspace of %zu bytes needed, %zu given (query tfx_filterbank_workspace_bytes)", need,
                              workspace_bytes);
                    return TFX_EWORKSPACE;
                }
            }
            if (prec == TFX_PREC_F32) {
                if constexpr (sizeof(IO) == 4)
                    rc = run_group_kb<IO, float>(Kb, plans, bands, nb, g, seg, stream);
                else
                    rc = TFX_EINVAL;
            } else {
                rc = run_group_kb<IO, double>(Kb, plans, bands, nb, g, seg, stream);
            }
            if (rc != TFX_OK) return rc;
            first_launch = false;
        }
    }
    return TFX_OK;
}

}  // namespace
}  // namespace tfx

extern "C" {

size_t tfx_filterbank_workspace_bytes(int64_t C, int64_t T, int N, int Kb) {
    (void)T;
    if (C <= 0 || N <= 0 || Kb <= 0) return 0;
    // S > 1 only when C*S fits one wave of streams (at most SMs*8*16 with 2 lanes per stream).
    const int64_t streams = static_cast<int64_t>(tfx::sm_count()) * tfx::kMaxCtasPerSm * tfx::kWarps * tfx::kMaxSpw + 128;
    const int nb = N < 32 ? N : 32;
    const size_t stream_path = static_cast<size_t>(2 * Kb * nb) * static_cast<size_t>(streams) * 8 + 256;
    // channel-tile SUM path (bank_tile.cu): up to 8 sections, kOversub work items per resident warp
    const int64_t tile_streams = tfx::bank_tile_stream_capacity() * tfx::kOversub + C + 128;
    const size_t tile_path = tfx::bank_sum_tile_ok(N, Kb, C) ? tfx::kWsHeader + static_cast<size_t>(2 * N * Kb) * static_cast<size_t>(tile_streams) * 8 + 256 : 0;
    const size_t stack_path = static_cast<size_t>(2 * Kb * nb) * static_cast<size_t>(tfx::bank_stack_max_streams() + C + 128) * 8 + 256;
    return std::max(std::max(stream_path, tile_path), stack_path);
}

int tfx_filterbank_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t ldb,
                       const double *sos_host, int N, int Kb, int mode, double *state_x, double *state_y, uint32_t flags,
                       void *workspace, size_t workspace_bytes, void *stream) {
    return tfx::filterbank_device<float>(x, y, C, T, ldx, ldy, ldb, sos_host, N, Kb, mode, state_x, state_y, flags,
                                         workspace, workspace_bytes, stream);
}

int tfx_filterbank_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t ldb,
                       const double *sos_host, int N, int Kb, int mode, double *state_x, double *state_y, uint32_t flags,
                       void *workspace, size_t workspace_bytes, void *stream) {
    return tfx::filterbank_device<double>(x, y, C, T, ldx, ldy, ldb, sos_host, N, Kb, mode, state_x, state_y, flags,
                                          workspace, workspace_bytes, stream);
}

}  // extern "C"
