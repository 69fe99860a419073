// sos_plan.h -- host-side analysis of an SOS cascade, done once per coefficient set and
// cached: (1) how fast the cascade forgets its state (-> how many warm-up samples a time
// segment needs before its state is exact to a given tolerance), (2) whether a float32
// recurrence is accurate enough (TFX_PREC_AUTO).
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

namespace tfx {

struct SosSection {
    double b0, b1, b2, a1, a2;
};

// A group of <= TFX_SOS_MAX_FUSED consecutive sections that one kernel launch fuses.
struct SosPass {
    int k0 = 0;  // first section
    int k = 0;   // number of sections
    // Smallest n with ||A^n||_inf <= tol / growth for the three tolerances used by the
    // kernels (A = DF2T state-transition matrix of these sections).  -1: does not decay
    // (unstable / marginally stable) -> never split this cascade in time.
    int64_t warm_f32 = -1;     // tol 2^-30 : float32 recurrence
    int64_t warm_f64_io32 = -1;  // tol 2^-42 : float64 recurrence, float32 I/O
    int64_t warm_f64_io64 = -1;  // tol 2^-62 : float64 recurrence, float64 I/O
};

struct SosPlan {
    int K = 0;
    std::vector<SosSection> sec;
    std::vector<SosPass> passes;
    int auto_prec = 0;        // TFX_PREC_F32 or TFX_PREC_F64
    double probe_rel_err = 0;  // max|y_f32 - y_f64| / max|y_f64| on the probe signal
    // When auto_prec == F64: the smallest set of sections (bit k = section k) that must run the
    // float64 recurrence for the probe error to fall under the bound; the others may stay
    // float32 (mixed-precision kernel).  All ones when no proper subset suffices.
    uint64_t mixed_mask = 0;
    double mixed_rel_err = 0;
};

// Cached by the coefficient bytes (thread-safe).  Returns nullptr and sets the error
// message when the coefficients are not finite or K is out of range.
std::shared_ptr<const SosPlan> get_sos_plan(const double *sos_host, int K);

// Time segmentation of one launch.
struct Segmentation {
    int64_t S = 1;     // segments per channel
    int64_t Lseg = 0;  // segment length in samples (multiple of 64 when S > 1)
    int64_t warm = 0;  // warm-up samples (multiple of 64; 0 when S == 1)
};
// `capacity` = number of streams the GPU holds in one wave (SMs * warps/SM * 32);
// `oversub` > 1 asks for that many work items per resident warp (dynamic scheduling).
Segmentation choose_segmentation(int64_t C, int64_t T, int64_t warm_needed, int64_t capacity, bool no_split,
                                 int oversub = 1);

}  // namespace tfx
