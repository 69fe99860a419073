// sos_tile.cu -- channel-tile variant of the fused SOS cascade: the main path when there are
// enough channels to fill the lanes (every BASELINE config).
//
// A warp owns 32 CONSECUTIVE CHANNELS over one time segment; its working set is a dense
// [32 channels x 64 samples] tile.  Compared with the stream-per-lane kernel
// (sos_cascade.cu, still used for few channels) all 32 lanes share the segment bounds, so
// every branch and loop count is warp-uniform, the row addresses are `base + r*ld` (no
// per-row offset tables, no ballots), and ragged edges exist only in the very last chunk of
// a channel.  Data path:
//   HBM --cp.async.cg 16 B (LDGSTS, 256 contiguous bytes per half-warp)--> shared memory,
//   stored with a 128-byte XOR swizzle (16-byte column v of row r lives at column v^(r&7))
//   so lane r reading "its" row with 128-bit LDS/STS is bank-conflict free without padding;
//   filtered in place; written back with coalesced 128-bit streaming stores.
// Round 1 also moved the same tiles with cp.async.bulk.tensor (TMA): on B200 that variant is
// capped near 4.6 TB/s by the per-SM TMA request rate on DRAM-missing 128-byte rows
// (profiles/r1_experiments.md section 3), the LDGSTS variant is not; the TMA kernel was removed.
//
// Work distribution: persistent warps pull (channel group, segment) items from a global
// counter; segment start states come from a warm-up launch (see sos_plan.cpp).
#include "sos_tile.cuh"

namespace tfx {
namespace {

template <typename IO, typename CT>
int launch_tile_any(const SosSection *sec, int k, const TileGeom &g, const Segmentation &seg, unsigned long long *counter,
                    cudaStream_t stream) {
    switch (k) {
        case 1: return launch_tile_k<IO, CT, 1>(sec, g, seg, counter, stream);
        case 2: return launch_tile_k<IO, CT, 2>(sec, g, seg, counter, stream);
        case 3: return launch_tile_k<IO, CT, 3>(sec, g, seg, counter, stream);
        case 4: return launch_tile_k<IO, CT, 4>(sec, g, seg, counter, stream);
        case 5: return launch_tile_k<IO, CT, 5>(sec, g, seg, counter, stream);
        case 6: return launch_tile_k<IO, CT, 6>(sec, g, seg, counter, stream);
        case 7: return launch_tile_k<IO, CT, 7>(sec, g, seg, counter, stream);
        case 8: return launch_tile_k<IO, CT, 8>(sec, g, seg, counter, stream);
        default: set_error("internal: pass with %d sections", k); return TFX_EINVAL;
    }
}

}  // namespace

int64_t tile_stream_capacity() { return static_cast<int64_t>(sm_count()) * kWarpsPerSm * 32; }

bool tile_path_ok(int64_t C) {
    const int64_t G = (C + 31) / 32;
    return C * 5 >= G * 32 * 4;  // >= 80 % of the lanes carry a channel
}

template <typename IO, typename CT>
int launch_tile_pass(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                     const Segmentation &seg, void *ws_base, double *state_x, double *state_y, cudaStream_t stream) {
    TileGeom g{};
    g.x = x;
    g.y = y;
    g.ldx = ldx;
    g.ldy = ldy;
    g.C = C;
    g.T = T;
    g.S = seg.S;
    g.Lseg = seg.Lseg;
    g.G = (C + 31) / 32;
    g.ws = ws_base ? static_cast<unsigned char *>(ws_base) + kWsHeader : nullptr;
    g.ws_stride = C * seg.S;
    g.state_x = state_x;
    g.state_y = state_y;
    const size_t esz = sizeof(IO);
    g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && ((ldx * esz) % 16 == 0) &&
               ((ldy * esz) % 16 == 0);
    return launch_tile_any<IO, CT>(sec, k, g, seg, static_cast<unsigned long long *>(ws_base), stream);
}

int launch_tile_pass_mixed(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                           unsigned f64_mask, const Segmentation &seg, void *ws_base, double *state_x, double *state_y,
                           cudaStream_t stream) {
    TileGeom g{};
    g.x = x;
    g.y = y;
    g.ldx = ldx;
    g.ldy = ldy;
    g.C = C;
    g.T = T;
    g.S = seg.S;
    g.Lseg = seg.Lseg;
    g.G = (C + 31) / 32;
    g.ws = ws_base ? static_cast<unsigned char *>(ws_base) + kWsHeader : nullptr;
    g.ws_stride = C * seg.S;
    g.state_x = state_x;
    g.state_y = state_y;
    g.f64_mask = f64_mask;
    g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && ((ldx * 4) % 16 == 0) &&
               ((ldy * 4) % 16 == 0);
    if (k > 4) return launch_tile_pass_mixed_long(x, y, C, T, ldx, ldy, sec, k, f64_mask, seg, ws_base, state_x, state_y, stream);
    unsigned long long *counter = static_cast<unsigned long long *>(ws_base);
#define TFX_MIXED_CASE(KK, MM) \
    if (k == KK && f64_mask == MM) return launch_tile_k<float, MixedF<MM>, KK>(sec, g, seg, counter, stream);
    TFX_MIXED_CASE(2, 1u) TFX_MIXED_CASE(2, 2u)
    TFX_MIXED_CASE(3, 1u) TFX_MIXED_CASE(3, 2u) TFX_MIXED_CASE(3, 3u) TFX_MIXED_CASE(3, 4u) TFX_MIXED_CASE(3, 5u) TFX_MIXED_CASE(3, 6u)
    TFX_MIXED_CASE(4, 1u) TFX_MIXED_CASE(4, 2u) TFX_MIXED_CASE(4, 3u) TFX_MIXED_CASE(4, 4u) TFX_MIXED_CASE(4, 5u) TFX_MIXED_CASE(4, 6u)
    TFX_MIXED_CASE(4, 7u) TFX_MIXED_CASE(4, 8u) TFX_MIXED_CASE(4, 9u) TFX_MIXED_CASE(4, 10u) TFX_MIXED_CASE(4, 11u)
    TFX_MIXED_CASE(4, 12u) TFX_MIXED_CASE(4, 13u) TFX_MIXED_CASE(4, 14u)
#undef TFX_MIXED_CASE
    set_error("internal: no mixed-precision kernel for K=%d mask=%u", k, f64_mask);
    return TFX_EINVAL;
}

bool tile_mixed_supported(int k, unsigned f64_mask) {
    if (k < 2 || k > 8 || f64_mask == 0u || f64_mask >= (1u << k) - 1u) return false;
    if (k <= 4) return true;                                  // every proper mask is instantiated (above)
    if ((f64_mask & (f64_mask - 1u)) == 0u) return true;      // a single section
    return f64_mask == 3u;                                    // the first two (sos_tile_mixed.cu)
}

// Smallest instantiated mask that contains `f64_mask` (0: none short of all sections -> run the float64 kernel).
unsigned tile_mixed_cover(int k, unsigned f64_mask) {
    if (tile_mixed_supported(k, f64_mask)) return f64_mask;
    if (k < 5 || k > 8 || f64_mask == 0u) return 0u;
    return (f64_mask & ~3u) == 0u ? 3u : 0u;                  // sections 0 and / or 1 -> the first two; anything else: float64 kernel
}

template int launch_tile_pass<float, float>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int,
                                            const Segmentation &, void *, double *, double *, cudaStream_t);
template int launch_tile_pass<float, double>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int,
                                             const Segmentation &, void *, double *, double *, cudaStream_t);
template int launch_tile_pass<double, double>(const double *, double *, int64_t, int64_t, int64_t, int64_t, const SosSection *,
                                              int, const Segmentation &, void *, double *, double *, cudaStream_t);

}  // namespace tfx
