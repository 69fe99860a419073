// bank_tile.cu -- SUM filterbank (`f1 + f2 + ...`) on the channel-tile kernel.
//
// Replaces ParallelFilterCombination.forward (filter/__base.py:1019-1026: N native calls,
// N temporaries, N adds) for banks of up to 8 sections in total: the N branches become the
// `Par` topology of sos_tile.cuh, i.e. ONE read of x and ONE write of y per sample
// (8 B per channel-sample, the same traffic as a single cascade), lanes = 32 consecutive
// channels, every branch's DF2T state in registers, coefficients in the constant bank, and
// N independent dependency chains per lane instead of one.  Larger banks, float64 I/O and
// few-channel inputs stay on bank_stream_kernel (filterbank.cu).
#include "bank_tile.h"
#include "sos_tile.cuh"

namespace tfx {
namespace {

template <typename CT>
int launch_par(const SosSection *sec, int N, int Kb, const TileGeom &g, const Segmentation &seg, unsigned long long *counter,
               cudaStream_t stream) {
#define TFX_PAR_CASE(NN, KK) \
    if (N == NN && Kb == KK) return launch_tile_k<float, Par<CT, KK>, NN * KK>(sec, g, seg, counter, stream);
    TFX_PAR_CASE(2, 1) TFX_PAR_CASE(3, 1) TFX_PAR_CASE(4, 1) TFX_PAR_CASE(5, 1) TFX_PAR_CASE(6, 1) TFX_PAR_CASE(7, 1) TFX_PAR_CASE(8, 1)
    TFX_PAR_CASE(2, 2) TFX_PAR_CASE(3, 2) TFX_PAR_CASE(4, 2)
#undef TFX_PAR_CASE
    set_error("internal: no parallel-bank tile kernel for N=%d Kb=%d", N, Kb);
    return TFX_EINVAL;
}

}  // namespace

bool bank_sum_tile_ok(int N, int Kb, int64_t C) {
    const bool shape = (Kb == 1 && N >= 2 && N <= 8) || (Kb == 2 && N >= 2 && N <= 4);
    const int64_t G = (C + 31) / 32;
    return shape && C * 5 >= G * 32 * 4;  // >= 80 % of the lanes carry a channel (as tile_path_ok)
}

int64_t bank_tile_stream_capacity() { return static_cast<int64_t>(sm_count()) * kWarpsPerSm * 32; }

template <typename CT>
int launch_bank_sum_tile(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int N, int Kb,
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
    g.vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && ((ldx * 4) % 16 == 0) &&
               ((ldy * 4) % 16 == 0);
    return launch_par<CT>(sec, N, Kb, g, seg, static_cast<unsigned long long *>(ws_base), stream);
}

template int launch_bank_sum_tile<float>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int, int,
                                         const Segmentation &, void *, double *, double *, cudaStream_t);
template int launch_bank_sum_tile<double>(const float *, float *, int64_t, int64_t, int64_t, int64_t, const SosSection *, int, int,
                                          const Segmentation &, void *, double *, double *, cudaStream_t);

}  // namespace tfx
