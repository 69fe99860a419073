// bank_tile_mixed.cu -- SUM filterbank on the channel-tile kernel, mixed precision: the first n64
// branches run the float64 recurrence, the others float32 (`ParMixed` topology of sos_tile.cuh).
// Picked by TFX_PREC_AUTO when the branches whose float32 recurrence fails the accuracy probe
// (sos_plan.cpp) form a prefix of the bank -- a bank listed by rising centre frequency.
#include "bank_tile.h"
#include "sos_tile.cuh"

namespace tfx {

bool bank_sum_mixed_ok(int N, int Kb, int n64) {
    return n64 >= 1 && n64 < N && ((Kb == 1 && N >= 2 && N <= 8) || (Kb == 2 && N >= 2 && N <= 4));
}

int launch_bank_sum_tile_mixed(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int N,
                               int Kb, int n64, const Segmentation &seg, void *ws_base, double *state_x, double *state_y,
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
    g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && ((ldx * 4) % 16 == 0) &&
               ((ldy * 4) % 16 == 0);
    unsigned long long *counter = static_cast<unsigned long long *>(ws_base);
#define TFX_PM_CASE(NN, KK, N64) \
    if (N == NN && Kb == KK && n64 == N64) return launch_tile_k<float, ParMixed<N64, KK>, NN * KK>(sec, g, seg, counter, stream);
    TFX_PM_CASE(2, 1, 1)
    TFX_PM_CASE(3, 1, 1) TFX_PM_CASE(3, 1, 2)
    TFX_PM_CASE(4, 1, 1) TFX_PM_CASE(4, 1, 2) TFX_PM_CASE(4, 1, 3)
    TFX_PM_CASE(5, 1, 1) TFX_PM_CASE(5, 1, 2) TFX_PM_CASE(5, 1, 3) TFX_PM_CASE(5, 1, 4)
    TFX_PM_CASE(6, 1, 1) TFX_PM_CASE(6, 1, 2) TFX_PM_CASE(6, 1, 3) TFX_PM_CASE(6, 1, 4) TFX_PM_CASE(6, 1, 5)
    TFX_PM_CASE(7, 1, 1) TFX_PM_CASE(7, 1, 2) TFX_PM_CASE(7, 1, 3) TFX_PM_CASE(7, 1, 4) TFX_PM_CASE(7, 1, 5) TFX_PM_CASE(7, 1, 6)
    TFX_PM_CASE(8, 1, 1) TFX_PM_CASE(8, 1, 2) TFX_PM_CASE(8, 1, 3) TFX_PM_CASE(8, 1, 4) TFX_PM_CASE(8, 1, 5) TFX_PM_CASE(8, 1, 6)
    TFX_PM_CASE(8, 1, 7)
    TFX_PM_CASE(2, 2, 1)
    TFX_PM_CASE(3, 2, 1) TFX_PM_CASE(3, 2, 2)
    TFX_PM_CASE(4, 2, 1) TFX_PM_CASE(4, 2, 2) TFX_PM_CASE(4, 2, 3)
#undef TFX_PM_CASE
    set_error("internal: no mixed parallel-bank tile kernel for N=%d Kb=%d n64=%d", N, Kb, n64);
    return TFX_EINVAL;
}

}  // namespace tfx
