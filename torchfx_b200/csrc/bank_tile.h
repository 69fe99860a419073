// bank_tile.h -- entry points of the channel-tile filterbank kernels (bank_tile.cu, bank_stack.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "sos_plan.h"

namespace tfx {

// SUM bank on the channel-tile cascade kernel (parallel topology): float32 I/O, N * Kb <= 8.
bool bank_sum_tile_ok(int N, int Kb, int64_t C);
int64_t bank_tile_stream_capacity();  // lanes resident in one wave
// sec: N * Kb sections, branch-major.  CT = float or double (recurrence precision).
template <typename CT>
int launch_bank_sum_tile(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int N, int Kb,
                         const Segmentation &seg, void *ws_base, double *state_x, double *state_y, cudaStream_t stream);

// Mixed precision (bank_tile_mixed.cu): branches [0, n64) on the float64 recurrence, the rest float32.
bool bank_sum_mixed_ok(int N, int Kb, int n64);
int launch_bank_sum_tile_mixed(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int N,
                               int Kb, int n64, const Segmentation &seg, void *ws_base, double *state_x, double *state_y,
                               cudaStream_t stream);

// STACK bank with lanes = channels (bank_stack.cu): float32 I/O, <= 32 bands per launch, per-band
// precision (bit b of f64_mask) and per-band warm-up lengths (warm_b[b] < 0: band does not decay).
bool bank_stack_tile_ok(int N, int Kb, int64_t C);
int64_t bank_stack_max_streams();  // upper bound of C * S (workspace sizing)
// False when this shape would leave most SMs idle (few channel groups x few time segments, even with the bands
// split over 4 CTAs): the band-per-lane kernel then fills the GPU better.  warm_max: longest band warm-up (< 0: none decays).
bool bank_stack_worthwhile(int64_t C, int64_t T, int nb, int Kb, int64_t warm_max, bool sum, bool no_split);
// sum = true: the bands are added into y[C, T] (`f1 + f2 + ...` with more than 8 sections in total) instead of stacked.
int launch_bank_stack(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int64_t ldb, const SosSection *sec,
                      const int *band_id, const int64_t *warm_b, uint32_t f64_mask, int nb, int Kb, bool sum, bool no_split,
                      void *workspace, size_t workspace_bytes, double *state_x, double *state_y, cudaStream_t stream);

}  // namespace tfx
