// sos_kernels.h -- entry points of the two cascade kernels (sos_tma.cu, sos_cascade.cu).
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

// Launch geometry shared by the cp.async stream kernels (sos_cascade.cu, sos_packed.cu).
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


// Packed-pair kernel (sos_packed.cu): float32 only, two streams per thread on FFMA2.
int64_t packed_stream_capacity();
int launch_packed_pass(const SosSection *sec, int k, Geom g, const Segmentation &seg, cudaStream_t stream);

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

// TMA-tiled kernel (sos_tma.cu): lanes = 32 consecutive channels.
bool tma_path_ok(const void *x, const void *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int elem_bytes);
int64_t tma_stream_capacity();  // streams (lanes) resident in one wave
template <typename IO, typename CT>
int launch_tma_pass(const IO *x, IO *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, const SosSection *sec, int k,
                    const Segmentation &seg, void *ws, double *state_x, double *state_y, cudaStream_t stream);

}  // namespace tfx
