// sos_kernels.h -- entry points of the two cascade kernels (sos_tma.cu, sos_cascade.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "sos_plan.h"

namespace tfx {

// TMA-tiled kernel (sos_tma.cu): lanes = 32 consecutive channels.
bool tma_path_ok(const void *x, const void *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int elem_bytes);
int64_t tma_stream_capacity();  // streams (lanes) resident in one wave
template <typename IO, typename CT>
int launch_tma_pass(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                    const Segmentation &seg, void *ws, double *state_x, double *state_y, cudaStream_t stream);

}  // namespace tfx
