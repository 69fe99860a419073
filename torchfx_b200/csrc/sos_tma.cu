// sos_tma.cu -- TMA-tiled variant of the fused SOS cascade (the main path on B200).
//
// Same decomposition as sos_cascade.cu (one thread = one stream = channel x time segment,
// whole cascade state in registers), different data movement:
//
//  * A warp owns 32 CONSECUTIVE CHANNELS over one time segment.  Its working set is a 2-D
//    tile [32 channels x 64 samples] of the row-major [C, T] signal, which is exactly what
//    the Tensor Memory Accelerator moves: one elected lane issues two
//    cp.async.bulk.tensor.2d loads (2 x [32 rows x 128 B], SWIZZLE_128B) per chunk and two
//    bulk tensor stores for the result -- no per-row address arithmetic, no LSU traffic for
//    the copies, hardware clipping / zero fill at the end of the signal and past the last
//    channel (no ragged-edge code on the memory path).
//  * The 128-byte swizzle places the 16-byte column v of row r at column v ^ (r & 7), so
//    lane r reading "its" row with 128-bit LDS/STS hits 8 distinct bank groups per quarter
//    warp: conflict-free without padding.
//  * Per-warp ring of kStages tiles guarded by mbarriers (complete_tx); stores drain
//    through cp.async.bulk commit/wait_group.  Warps never synchronise with each other.
//
// Requirements (else the generic cp.async kernel in sos_cascade.cu runs): 16-byte aligned
// base pointers and row pitches, T < 2^31, enough channels to fill the lanes.
#include <algorithm>
#include <cstdint>
#include <mutex>

#include "common.cuh"
#include "sos_kernels.h"
#include "stream_common.cuh"
#include "tma.cuh"

namespace tfx {

// ---- host: tensor-map encoder through the driver entry point ----------------------------------
namespace {
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = reinterpret_cast<EncodeFn>(fn);
    else
        (void)cudaGetLastError();
}
}  // namespace

bool tma_available() {
    std::call_once(g_encode_once, resolve_encode);
    return g_encode != nullptr;
}

int encode_tile_map_2d(CUtensorMap *map, const void *base, int elem_bytes, uint64_t cols, uint64_t rows,
                       uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows) {
    if (!tma_available()) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return TFX_ECUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {row_pitch_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const CUresult r = g_encode(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu pitch=%llu)", static_cast<int>(r),
                  (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_pitch_bytes);
        return TFX_ECUDA;
    }
    return TFX_OK;
}

namespace {

#ifndef TFX_TMA_STAGES
#define TFX_TMA_STAGES 4
#endif
#ifndef TFX_TMA_WARPS
#define TFX_TMA_WARPS 1
#endif
constexpr int kStages = TFX_TMA_STAGES;  // tiles per warp ring: 1 computing + kAhead in flight + kDrain draining their stores
#ifndef TFX_TMA_AHEAD
#define TFX_TMA_AHEAD (TFX_TMA_STAGES - 2)
#endif
constexpr int kAhead = TFX_TMA_AHEAD;  // prefetch distance in chunks
// The tile refilled at iteration i held chunk i + kAhead - kStages; the stores of the
// kStages - kAhead - 1 chunks issued after it may still be reading shared memory.
constexpr int kDrain = kStages - kAhead - 1;
static_assert(kAhead >= 1 && kDrain >= 1, "ring needs one tile computing, kAhead in flight and >= 1 draining");
constexpr int kWarps = TFX_TMA_WARPS;    // warps per CTA (warps are fully independent)
constexpr int kTileBytes = 32 * 256;     // [32 channels x 256 B] = two swizzled [32 x 128 B] boxes
constexpr int kWarpSmem = kStages * kTileBytes + 128;  // + mbarriers
constexpr int kCtaSmem = kWarps * kWarpSmem + 1024;    // + slack to align the ring to 1024 B
constexpr int kCtasPerSm = std::min(32, kSmemPerSm / (kCtaSmem + 1024));
constexpr int kWarpsPerSm = kCtasPerSm * kWarps;

struct TmaGeom {
    int64_t C, T;
    int64_t S, Lseg, warm;
    int64_t nwarps;  // work items in this launch: G * S (main) or G * (S - 1) (warm-up)
    int64_t G;       // channel groups of 32
    void *ws;        // [2K][C * S]
    int64_t ws_stride;
    double *state_x;
    double *state_y;
};

template <typename IO>
__device__ __forceinline__ int elem_offset(int lane, int e) {
    // byte offset of element e (0..CHUNK) of row `lane` inside a swizzled tile
    constexpr int EPB = 128 / sizeof(IO);  // elements per 128-byte box row
    constexpr int EPV = 16 / sizeof(IO);   // elements per 16-byte column
    const int sub = e / EPB, ee = e % EPB;
    return sub * 4096 + lane * 128 + (((ee / EPV) ^ (lane & 7)) << 4) + (ee % EPV) * static_cast<int>(sizeof(IO));
}

template <typename IO, typename CT, int K>
__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm)
sos_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
               const __grid_constant__ SosCoef<CT, K> cf, const __grid_constant__ SosCoefD<K> cd,
               const __grid_constant__ TmaGeom g) {
    using Tr = IoTraits<IO>;
    using Vec = typename Tr::Vec;
    constexpr int CH = 256 / sizeof(IO);  // samples per chunk
    constexpr int UV = K <= 2 ? 16 : (K <= 4 ? 8 : 4);

    // plain pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *cta_base = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char *ring = cta_base + warp * (kStages * kTileBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(cta_base + kWarps * (kStages * kTileBytes) + warp * 128);

    const int64_t w = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
    if (w >= g.nwarps) return;  // whole warp exits together
    const bool warm_pass = g.warm > 0;
    int64_t grp, j, n0, n1;
    if (warm_pass) {
        const int64_t sm1 = g.S - 1;
        grp = w / sm1;
        j = w - grp * sm1 + 1;
        n1 = j * g.Lseg;
        n0 = max(n1 - g.warm, static_cast<int64_t>(0));
    } else {
        grp = w / g.S;
        j = w - grp * g.S;
        n0 = j * g.Lseg;
        n1 = min(g.T, n0 + g.Lseg);
    }
    const int64_t c = grp * 32 + lane;
    const bool live = c < g.C;
    const bool from_true_state = n0 == 0;
    const bool do_tail = !warm_pass && (j == g.S - 1) && g.state_x != nullptr;  // warp-uniform
    const int64_t len = n1 - n0;
    const int64_t nch = (len + CH - 1) / CH;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        fence_proxy_async_smem();
    }
    __syncwarp();

    // ---- start state (DF2T) ----------------------------------------------------------------
    CT s1[K], s2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s1[k] = CT(0);
        s2[k] = CT(0);
    }
    if (live) {
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
        } else if (!warm_pass) {
            const CT *wsp = static_cast<const CT *>(g.ws) + (c * g.S + j);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                s1[k] = wsp[(2 * k) * g.ws_stride];
                s2[k] = wsp[(2 * k + 1) * g.ws_stride];
            }
        }
    }

    const int32_t row0 = static_cast<int32_t>(grp * 32);
    auto issue_load = [&](int64_t i, int stage) {  // lane 0 only
        unsigned char *tile = ring + stage * kTileBytes;
        const int32_t col = static_cast<int32_t>(n0 + i * CH);
        mbar_arrive_expect_tx(&bars[stage], kTileBytes);
        tma_load_2d(tile, &map_x, col, row0, &bars[stage]);
        tma_load_2d(tile + 4096, &map_x, col + CH / 2, row0, &bars[stage]);
    };

    if (lane == 0) {
        prefetch_tensormap(&map_x);
        if (!warm_pass) prefetch_tensormap(&map_y);
#pragma unroll
        for (int s = 0; s < kAhead; ++s)
            if (s < nch) issue_load(s, s);
    }

    // DF1 history of every section, only maintained over the channel's last two chunks
    CT hx[K][2], hy[K][2];
    if (do_tail) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            hx[k][0] = hx[k][1] = hy[k][0] = hy[k][1] = CT(0);
            if (live && from_true_state) {  // consulted only when fewer than two samples are filtered (then S == 1)
                const int64_t o = (static_cast<int64_t>(k) * g.C + c) * 2;
                hx[k][0] = static_cast<CT>(g.state_x[o]);
                hx[k][1] = static_cast<CT>(g.state_x[o + 1]);
                hy[k][0] = static_cast<CT>(g.state_y[o]);
                hy[k][1] = static_cast<CT>(g.state_y[o + 1]);
            }
        }
    }

    int stage = 0;
    uint32_t parity = 0;
    for (int64_t i = 0; i < nch; ++i) {
        // keep the ring full: chunk i+kAhead goes into the tile that held chunk i+kAhead-kStages,
        // reusable once that chunk's store has finished reading it (the kDrain newer stores may still be in flight)
        if (lane == 0 && i + kAhead < nch) {
            tma_store_wait_read<kDrain>();
            int st = stage + kAhead;
            if (st >= kStages) st -= kStages;
            issue_load(i + kAhead, st);
        }
        mbar_wait(&bars[stage], parity);
        unsigned char *tile = ring + stage * kTileBytes;
        const int64_t base = i * CH;
        const int cnt = static_cast<int>(min(len - base, static_cast<int64_t>(CH)));
        const bool tracked = do_tail && i >= nch - 2;

        if (live) {
            if (cnt == CH && !tracked) {
                auto vec_at = [&](int v) { return reinterpret_cast<Vec *>(tile + (v >> 3) * 4096 + lane * 128 + (((v & 7) ^ (lane & 7)) << 4)); };
                Vec a = *vec_at(0);
#pragma unroll UV
                for (int v = 0; v < 16; ++v) {
                    const Vec nxt = *vec_at((v + 1) & 15);  // fetched before the in-place store below
                    filter_vec<CT, K>(cf, s1, s2, a);
                    *vec_at(v) = a;
                    a = nxt;
                }
            } else if (!tracked) {
                for (int e = 0; e < cnt; ++e) {
                    IO *p = reinterpret_cast<IO *>(tile + elem_offset<IO>(lane, e));
                    *p = static_cast<IO>(sos_step<CT, K>(cf, s1, s2, static_cast<CT>(*p)));
                }
            } else {
                for (int e = 0; e < cnt; ++e) {
                    IO *p = reinterpret_cast<IO *>(tile + elem_offset<IO>(lane, e));
                    CT v = static_cast<CT>(*p);
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
                    *p = static_cast<IO>(v);
                }
            }
        }

        if (!warm_pass) {
            fence_proxy_async_smem();  // my STS must be visible to the TMA store
            __syncwarp();
            if (lane == 0) {
                const int32_t col = static_cast<int32_t>(n0 + base);
                tma_store_2d(&map_y, col, row0, tile);
                tma_store_2d(&map_y, col + CH / 2, row0, tile + 4096);
                tma_store_commit();
            }
        } else {
            __syncwarp();
        }
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1;
        }
    }
    if (lane == 0 && !warm_pass) tma_store_wait_all<0>();

    if (!live) return;
    if (warm_pass) {
        CT *wsp = static_cast<CT *>(g.ws) + (c * g.S + j);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            wsp[(2 * k) * g.ws_stride] = s1[k];
            wsp[(2 * k + 1) * g.ws_stride] = s2[k];
        }
        return;
    }
    if (do_tail) {
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

template <typename IO, typename CT, int K>
int launch_tma_k(const SosSection *sec, const CUtensorMap &mx, const CUtensorMap &my, TmaGeom g, const Segmentation &seg,
                 cudaStream_t stream) {
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
    auto kern = sos_tma_kernel<IO, CT, K>;
    TFX_ENSURE_SMEM(kern, kCtaSmem);
    if (seg.S > 1) {
        TmaGeom gw = g;
        gw.warm = seg.warm;
        gw.nwarps = g.G * (seg.S - 1);
        const int64_t grid = (gw.nwarps + kWarps - 1) / kWarps;
        kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(mx, mx, cf, cd, gw);
        TFX_CHECK_LAUNCH("sos_tma_kernel(warm-up)");
    }
    g.warm = 0;
    g.nwarps = g.G * seg.S;
    const int64_t grid = (g.nwarps + kWarps - 1) / kWarps;
    kern<<<static_cast<unsigned>(grid), kWarps * 32, kCtaSmem, stream>>>(mx, my, cf, cd, g);
    TFX_CHECK_LAUNCH("sos_tma_kernel");
    return TFX_OK;
}

template <typename IO, typename CT>
int launch_tma_any(const SosSection *sec, int k, const CUtensorMap &mx, const CUtensorMap &my, const TmaGeom &g,
                   const Segmentation &seg, cudaStream_t stream) {
    switch (k) {
        case 1: return launch_tma_k<IO, CT, 1>(sec, mx, my, g, seg, stream);
        case 2: return launch_tma_k<IO, CT, 2>(sec, mx, my, g, seg, stream);
        case 3: return launch_tma_k<IO, CT, 3>(sec, mx, my, g, seg, stream);
        case 4: return launch_tma_k<IO, CT, 4>(sec, mx, my, g, seg, stream);
        case 5: return launch_tma_k<IO, CT, 5>(sec, mx, my, g, seg, stream);
        case 6: return launch_tma_k<IO, CT, 6>(sec, mx, my, g, seg, stream);
        case 7: return launch_tma_k<IO, CT, 7>(sec, mx, my, g, seg, stream);
        case 8: return launch_tma_k<IO, CT, 8>(sec, mx, my, g, seg, stream);
        default: set_error("internal: pass with %d sections", k); return TFX_EINVAL;
    }
}

}  // namespace

int64_t tma_stream_capacity() { return static_cast<int64_t>(sm_count()) * kWarpsPerSm * 32; }

bool tma_path_ok(const void *x, const void *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int elem_bytes) {
    if (!tma_available()) return false;
    if (T >= (int64_t(1) << 31) - 512 || C >= (int64_t(1) << 31) - 64) return false;
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(y) % 16) return false;
    if ((ldx * elem_bytes) % 16 || (ldy * elem_bytes) % 16) return false;
    // lanes are channels: require the last group of 32 to be reasonably full
    const int64_t G = (C + 31) / 32;
    return C * 5 >= G * 32 * 4;  // >= 80 % lane utilisation
}

template <typename IO, typename CT>
int launch_tma_pass(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                    const Segmentation &seg, void *ws, double *state_x, double *state_y, cudaStream_t stream) {
    CUtensorMap mx, my;
    constexpr uint32_t box_cols = 128 / sizeof(IO);
    int rc = encode_tile_map_2d(&mx, x, sizeof(IO), static_cast<uint64_t>(T), static_cast<uint64_t>(C),
                                static_cast<uint64_t>(ldx) * sizeof(IO), box_cols, 32);
    if (rc != TFX_OK) return rc;
    rc = encode_tile_map_2d(&my, y, sizeof(IO), static_cast<uint64_t>(T), static_cast<uint64_t>(C),
                            static_cast<uint64_t>(ldy) * sizeof(IO), box_cols, 32);
    if (rc != TFX_OK) return rc;
    TmaGeom g{};
    g.C = C;
    g.T = T;
    g.S = seg.S;
    g.Lseg = seg.Lseg;
    g.G = (C + 31) / 32;
    g.ws = ws;
    g.ws_stride = C * seg.S;
    g.state_x = state_x;
    g.state_y = state_y;
    return launch_tma_any<IO, CT>(sec, k, mx, my, g, seg, stream);
}

template int launch_tma_pass<float, float>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int,
                                           const Segmentation &, void *, double *, double *, cudaStream_t);
template int launch_tma_pass<float, double>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int,
                                            const Segmentation &, void *, double *, double *, cudaStream_t);
template int launch_tma_pass<double, double>(const double *, double *, int64_t, int64_t, int64_t, int64_t, const SosSection *,
                                             int, const Segmentation &, void *, double *, double *, cudaStream_t);

}  // namespace tfx
