// fir_ols16k.h -- entry points of the persistent overlap-save FIR kernel (fir_ols16k.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace tfx {

// Workspace: completion counters, twiddle tables, the taps' spectra H[P] and the L2-resident spectra ring /
// product rows of the G open channel pairs.  Depends on the tap count only.
size_t fir_ols16k_workspace_bytes(int64_t K);

// A plan holds what depends on the impulse response only (twiddle tables + the taps' spectra H[P]) in a caller-owned
// device buffer, so that chunked callers do not recompute it on every call.
size_t fir_ols16k_plan_bytes(int64_t K);
int fir_ols16k_plan_init(const float *taps, int64_t K, void *plan, size_t plan_bytes, cudaStream_t stream);

// y[c, n] = sum_j taps[j] * x[c, n - j], zero history; float32, not in place.  plan == NULL: the twiddles and the taps'
// spectra are computed from `taps` into the workspace by this call; else `taps` is not read.
int launch_fir_ols16k(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const float *taps, const void *plan,
                      int64_t K, void *workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace tfx
