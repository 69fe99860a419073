/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Plain-C CPU restatement of the arithmetic of the reference's native hot path
 * (matteospanio/torchfx @ 27e65b0, src/torchfx/_csrc/cpu/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; torchfx_b200/ never does.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   (a) golden vectors in tests/golden/ produced by the unmodified reference
 *       (oracle/make_golden.py imports /root/reference with oracle/_ref/torchfx_ext.so),
 *   (b) scipy.signal.sosfilt / lfilter, the oracle the reference's own tests use
 *       (tests/test_ops_dispatch.py:74-127, tests/test_fused.py:116-137).
 *
 * Every function cites the reference lines it follows.  All arithmetic is IEEE
 * double, evaluated in the reference's operation order and compiled WITHOUT
 * -ffast-math, so the oracle is deterministic (the reference itself is built with
 * -ffast-math, CMakeLists.txt:109, hence parity with it is to ~1 ulp of f64, not
 * bitwise).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- *
 * Single biquad, Direct Form 1, f64.
 * Follows cpu/iir_cpu.cpp:10-62 (biquad_forward_cpu):
 *   yn = b0*xn + b1*sx0 + b2*sx1 - a1*sy0 - a2*sy1       (iir_cpu.cpp:46)
 * state_x[c] = {x[n-1], x[n-2]}, state_y[c] = {y[n-1], y[n-2]}  (iir_cpu.cpp:39-42)
 * State is updated in place here (the reference clones and returns it,
 * iir_cpu.cpp:20-21,55-58 -- same values).
 * ------------------------------------------------------------------------- */
void oracle_biquad_df1_f64(const double *x, double *y, int64_t C, int64_t T,
                           double b0, double b1, double b2, double a1, double a2,
                           double *state_x /* [C,2] */, double *state_y /* [C,2] */)
{
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        double sx0 = state_x[2 * c + 0], sx1 = state_x[2 * c + 1];
        double sy0 = state_y[2 * c + 0], sy1 = state_y[2 * c + 1];
        const double *xc = x + c * T;
        double *yc = y + c * T;
        for (int64_t n = 0; n < T; ++n) {
            double xn = xc[n];
            double yn = b0 * xn + b1 * sx0 + b2 * sx1 - a1 * sy0 - a2 * sy1;
            yc[n] = yn;
            sx1 = sx0; sx0 = xn;
            sy1 = sy0; sy0 = yn;
        }
        state_x[2 * c + 0] = sx0; state_x[2 * c + 1] = sx1;
        state_y[2 * c + 0] = sy0; state_y[2 * c + 1] = sy1;
    }
}

/* ------------------------------------------------------------------------- *
 * K-section SOS cascade, Direct Form 1 per section, f64, fused over sections.
 * Follows cpu/iir_cpu.cpp:64-159 (sos_forward_cpu), hot loop :132-147:
 *   for c: for n: val = x[c][n]; for s: yn = b0*val + b1*sx0[s] + b2*sx1[s]
 *                                          - a1*sy0[s] - a2*sy1[s]; ...; val = yn
 * sos rows are [b0 b1 b2 a0 a1 a2] with a0 == 1 ignored (iir_cpu.cpp:84-86).
 * state_x / state_y are [K, C, 2] (iir_cpu.cpp:94-95,125-130,150-155).
 * ------------------------------------------------------------------------- */
void oracle_sos_df1_f64(const double *x, double *y, int64_t C, int64_t T,
                        const double *sos /* [K,6] */, int64_t K,
                        double *state_x /* [K,C,2] */, double *state_y /* [K,C,2] */)
{
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        double *sx0 = (double *)malloc(sizeof(double) * 4 * (size_t)(K > 0 ? K : 1));
        double *sx1 = sx0 + K, *sy0 = sx1 + K, *sy1 = sy0 + K;
        for (int64_t s = 0; s < K; ++s) {
            sx0[s] = state_x[(s * C + c) * 2 + 0];
            sx1[s] = state_x[(s * C + c) * 2 + 1];
            sy0[s] = state_y[(s * C + c) * 2 + 0];
            sy1[s] = state_y[(s * C + c) * 2 + 1];
        }
        const double *xc = x + c * T;
        double *yc = y + c * T;
        for (int64_t n = 0; n < T; ++n) {
            double val = xc[n];
            for (int64_t s = 0; s < K; ++s) {
                const double *co = sos + 6 * s;
                double yn = co[0] * val + co[1] * sx0[s] + co[2] * sx1[s]
                          - co[4] * sy0[s] - co[5] * sy1[s];
                sx1[s] = sx0[s]; sx0[s] = val;
                sy1[s] = sy0[s]; sy0[s] = yn;
                val = yn;
            }
            yc[n] = val;
        }
        for (int64_t s = 0; s < K; ++s) {
            state_x[(s * C + c) * 2 + 0] = sx0[s];
            state_x[(s * C + c) * 2 + 1] = sx1[s];
            state_y[(s * C + c) * 2 + 0] = sy0[s];
            state_y[(s * C + c) * 2 + 1] = sy1[s];
        }
        free(sx0);
    }
}

/* ------------------------------------------------------------------------- *
 * The reference's float32 contract: x (f32) -> f64 (_ops.py:142), f64 cascade,
 * result cast back to the input dtype (filter/iir.py:176).  Row strides in
 * elements so the oracle can be run on sub-blocks of a larger signal.
 * ------------------------------------------------------------------------- */
void oracle_sos_df1_f32io(const float *x, float *y, int64_t C, int64_t T,
                          int64_t ldx, int64_t ldy,
                          const double *sos, int64_t K,
                          double *state_x, double *state_y)
{
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        double *st = (double *)malloc(sizeof(double) * 4 * (size_t)(K > 0 ? K : 1));
        double *sx0 = st, *sx1 = sx0 + K, *sy0 = sx1 + K, *sy1 = sy0 + K;
        for (int64_t s = 0; s < K; ++s) {
            sx0[s] = state_x[(s * C + c) * 2 + 0];
            sx1[s] = state_x[(s * C + c) * 2 + 1];
            sy0[s] = state_y[(s * C + c) * 2 + 0];
            sy1[s] = state_y[(s * C + c) * 2 + 1];
        }
        const float *xc = x + c * ldx;
        float *yc = y + c * ldy;
        for (int64_t n = 0; n < T; ++n) {
            double val = (double)xc[n];
            for (int64_t s = 0; s < K; ++s) {
                const double *co = sos + 6 * s;
                double yn = co[0] * val + co[1] * sx0[s] + co[2] * sx1[s]
                          - co[4] * sy0[s] - co[5] * sy1[s];
                sx1[s] = sx0[s]; sx0[s] = val;
                sy1[s] = sy0[s]; sy0[s] = yn;
                val = yn;
            }
            yc[n] = (float)val;
        }
        for (int64_t s = 0; s < K; ++s) {
            state_x[(s * C + c) * 2 + 0] = sx0[s];
            state_x[(s * C + c) * 2 + 1] = sx1[s];
            state_y[(s * C + c) * 2 + 0] = sy0[s];
            state_y[(s * C + c) * 2 + 1] = sy1[s];
        }
        free(st);
    }
}

/* ------------------------------------------------------------------------- *
 * Delay line (the other native op).
 * Follows cpu/delay_cpu.cpp:17-43: out[n] = in[n] (n < D);
 * out[n] = in[n] + coeff*in[n-D] (n >= D), coeff = mix*decay computed in double
 * and, for f32, rounded once to float (delay_cpu.cpp:69-75).
 * The T <= D short-circuit (delay_cpu.cpp:62-64) returns the input unchanged.
 * ------------------------------------------------------------------------- */
void oracle_delay_line_f32(const float *x, float *y, int64_t C, int64_t T,
                           int64_t delay, double decay, double mix)
{
    if (T <= delay) { memcpy(y, x, sizeof(float) * (size_t)(C * T)); return; }
    const float coeff = (float)(mix * decay);
    for (int64_t c = 0; c < C; ++c) {
        const float *in = x + c * T; float *out = y + c * T;
        for (int64_t n = 0; n < delay; ++n) out[n] = in[n];
        for (int64_t n = delay; n < T; ++n) out[n] = in[n] + coeff * in[n - delay];
    }
}

void oracle_delay_line_f64(const double *x, double *y, int64_t C, int64_t T,
                           int64_t delay, double decay, double mix)
{
    if (T <= delay) { memcpy(y, x, sizeof(double) * (size_t)(C * T)); return; }
    const double coeff = mix * decay;
    for (int64_t c = 0; c < C; ++c) {
        const double *in = x + c * T; double *out = y + c * T;
        for (int64_t n = 0; n < delay; ++n) out[n] = in[n];
        for (int64_t n = delay; n < T; ++n) out[n] = in[n] + coeff * in[n - delay];
    }
}

/* ------------------------------------------------------------------------- *
 * Causal FIR, zero history, output length T:  y[n] = sum_j b[j] * x[n-j].
 * This is what FIR.forward computes (filter/fir.py:526-579: left-pad K-1 zeros,
 * cross-correlate with the flipped kernel; filter/_fftconv.py:107-141 evaluates the
 * same sum by overlap-save) -- equivalently F.conv1d(F.pad(x,(K-1,0)), b.flip()),
 * the oracle of the reference's tests/test_fftconv.py:65-77.  Accumulated in f64 and
 * rounded once, so it bounds BOTH float32 evaluation orders (FFT and direct).
 * ------------------------------------------------------------------------- */
void oracle_fir_causal_f32(const float *x, float *y, int64_t C, int64_t T,
                           const float *b, int64_t K)
{
#pragma omp parallel for schedule(static) if (C > 1)
    for (int64_t c = 0; c < C; ++c) {
        const float *xc = x + c * T; float *yc = y + c * T;
        for (int64_t n = 0; n < T; ++n) {
            double acc = 0.0;
            int64_t jmax = n < K - 1 ? n : K - 1;
            for (int64_t j = 0; j <= jmax; ++j) acc += (double)b[j] * (double)xc[n - j];
            yc[n] = (float)acc;
        }
    }
}
