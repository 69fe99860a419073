// sos_tile_mixed.cu -- mixed-precision instantiations of the channel-tile cascade kernel for 5..8 fused sections.
//
// TFX_PREC_AUTO runs a cascade on the float64 recurrence when the float32 one misses the 2e-6 probe bound
// (sos_plan.cpp), and for cascades of up to 4 sections every "float64 only in these sections" mask has its own
// instantiation (sos_tile.cu).  Longer chains usually fail the probe because of ONE float32-hostile section (a
// 20 Hz high-pass in front of an equaliser chain, a narrow notch): round 1 ran all of their sections in float64
// (K = 6: 466 against 694 Gsamples/s).  The mask is a template constant (a run-time per-section branch measured
// slower than float64 everywhere), so the family instantiated here is: any single section, and the first two.
// Anything else runs the float64 kernel: with three or more float64 sections the mixed instantiation no longer
// fits the register budget of the full-residency launch (measured, K = 7 with the first six sections in float64:
// 50 Gsamples/s against 83 for the all-float64 kernel).
#include "sos_tile.cuh"

namespace tfx {

int launch_tile_pass_mixed_long(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
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
    unsigned long long *counter = static_cast<unsigned long long *>(ws_base);
#define TFX_MIXED_CASE(KK, MM) \
    if (k == KK && f64_mask == MM) return launch_tile_k<float, MixedF<MM>, KK>(sec, g, seg, counter, stream);
    // single sections
    TFX_MIXED_CASE(5, 1u) TFX_MIXED_CASE(5, 2u) TFX_MIXED_CASE(5, 4u) TFX_MIXED_CASE(5, 8u) TFX_MIXED_CASE(5, 16u)
    TFX_MIXED_CASE(6, 1u) TFX_MIXED_CASE(6, 2u) TFX_MIXED_CASE(6, 4u) TFX_MIXED_CASE(6, 8u) TFX_MIXED_CASE(6, 16u) TFX_MIXED_CASE(6, 32u)
    TFX_MIXED_CASE(7, 1u) TFX_MIXED_CASE(7, 2u) TFX_MIXED_CASE(7, 4u) TFX_MIXED_CASE(7, 8u) TFX_MIXED_CASE(7, 16u) TFX_MIXED_CASE(7, 32u)
    TFX_MIXED_CASE(7, 64u)
    TFX_MIXED_CASE(8, 1u) TFX_MIXED_CASE(8, 2u) TFX_MIXED_CASE(8, 4u) TFX_MIXED_CASE(8, 8u) TFX_MIXED_CASE(8, 16u) TFX_MIXED_CASE(8, 32u)
    TFX_MIXED_CASE(8, 64u) TFX_MIXED_CASE(8, 128u)
    // the first two sections (a high-pass + a low notch in front of the chain)
    TFX_MIXED_CASE(5, 3u) TFX_MIXED_CASE(6, 3u) TFX_MIXED_CASE(7, 3u) TFX_MIXED_CASE(8, 3u)
#undef TFX_MIXED_CASE
    set_error("internal: no mixed-precision kernel for K=%d mask=%u", k, f64_mask);
    return TFX_EINVAL;
}

}  // namespace tfx
