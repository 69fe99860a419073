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

}  // namespace tfx
