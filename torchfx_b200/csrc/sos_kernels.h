// sos_kernels.h -- entry points of the cascade kernels (sos_cascade.cu: stream per lane; sos_tile.cu: lanes = channels).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "sos_plan.h"

namespace tfx {

constexpr size_t kWsHeader = 256;  // workspace header (work counter)
#ifndef TFX_OVERSUB
#define TFX_OVERSUB 16
#endif
constexpr int kOversub = TFX_OVERSUB;  // work items per resident warp when the signal is long enough

// Launch geometry of the stream-per-lane kernel (sos_cascade.cu).
struct Geom {
    const void *x;
    void *y;
    int64_t ldx, ldy, C, T;
    int64_t S;         // segments per channel
    int64_t Lseg;      // segment length
    int64_t warm;      // > 0: this launch is the warm-up pass over segments 1..S-1
    int64_t nstreams;  // streams in this launch
    void *ws_base;     // workspace start: [0, 256) work counter, then the states
    void *ws;          // [2K][C*S] segment start states (compute type)
    int64_t ws_stride;
    double *state_x;  // [K, C, 2] DF1 state of THIS pass's sections (or NULL)
    double *state_y;
    int vec_ok;  // all rows 16-byte aligned -> 128-bit global accesses allowed
    unsigned long long *counter;  // dynamic work distribution (NULL: one item per warp)
};


// Channel-tile kernel (sos_tile.cu): lanes = 32 consecutive channels, cp.async tiles.
bool tile_path_ok(int64_t C);
int64_t tile_stream_capacity();
template <typename IO, typename CT>
int launch_tile_pass(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                     const Segmentation &seg, void *ws_base, double *state_x, double *state_y, cudaStream_t stream);

// Mixed precision: float32 recurrence, float64 only in the sections of `f64_mask` (bit k = section k).
bool tile_mixed_supported(int k, unsigned f64_mask);
unsigned tile_mixed_cover(int k, unsigned f64_mask);  // smallest instantiated superset of the mask, 0 if none
int launch_tile_pass_mixed_long(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                                unsigned f64_mask, const Segmentation &seg, void *ws_base, double *state_x, double *state_y,
                                cudaStream_t stream);  // K = 5..8 (sos_tile_mixed.cu)
int launch_tile_pass_mixed(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                           unsigned f64_mask, const Segmentation &seg, void *ws_base, double *state_x, double *state_y,
                           cudaStream_t stream);

}  // namespace tfx
