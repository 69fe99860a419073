// filterbank.cu -- placeholder until the band-per-lane kernel lands (next commit).
#include "common.cuh"
extern "C" {
size_t tfx_filterbank_workspace_bytes(int64_t, int64_t, int, int) { return 0; }
int tfx_filterbank_f32(const float *, float *, int64_t, int64_t, int64_t, int64_t, int64_t, const double *, int, int, int,
                       double *, double *, uint32_t, void *, size_t, void *) {
    tfx::set_error("filterbank: not built yet");
    return TFX_EINVAL;
}
int tfx_filterbank_f64(const double *, double *, int64_t, int64_t, int64_t, int64_t, int64_t, const double *, int, int, int,
                       double *, double *, uint32_t, void *, size_t, void *) {
    tfx::set_error("filterbank: not built yet");
    return TFX_EINVAL;
}
}
