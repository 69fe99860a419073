// fir.cu -- placeholder until the FIR kernels land.
#include "common.cuh"
extern "C" {
size_t tfx_fir_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }
int tfx_fir_f32(const float *, float *, int64_t, int64_t, int64_t, int64_t, const float *, int64_t, int, void *, size_t,
                void *) {
    tfx::set_error("fir: not built yet");
    return TFX_EINVAL;
}
}
