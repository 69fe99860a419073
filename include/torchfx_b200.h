/*
 * torchfx_b200.h -- C ABI of the B200-native filter engine (libtorchfx_b200.so).
 *
 * This is the drop-in boundary for the reference's native layer, the pybind11 module
 * `torchfx.torchfx_ext` (reference: src/torchfx/_csrc/binding.cpp:83-96) and its Python
 * dispatch wrappers (src/torchfx/_ops.py:57-191).  Plain pointers and sizes only: no
 * torch / pybind types.  torchfx_b200/_native.py binds it with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add to src/torchfx/_ops.py.
 *
 * Conventions
 *  - Signals are row-major [C, T] with a row stride in ELEMENTS (ldx / ldy); [B, C, T]
 *    is flattened to [B*C, T] by the caller exactly as the reference does
 *    (src/torchfx/filter/iir.py:119-126).
 *  - `*_f32` / `*_f64` name the I/O element type.  Arithmetic type is chosen by
 *    `flags & TFX_PREC_MASK`.
 *  - SOS coefficients are HOST pointers to [K, 6] doubles, rows [b0 b1 b2 a0 a1 a2],
 *    a0 ignored (== 1), the layout of scipy.signal `output="sos"` and of the reference
 *    (src/torchfx/_csrc/cpu/iir_cpu.cpp:82-87).
 *  - Filter state is the reference's DF1 contract: state_x / state_y are [K, C, 2]
 *    doubles holding {v[n-1], v[n-2]} of every section's input / output
 *    (src/torchfx/_csrc/cpu/iir_cpu.cpp:39-42,125-130).  Device entry points take DEVICE
 *    state pointers, `_cpu` / `_host` entry points HOST pointers.  State is updated
 *    IN PLACE; NULL means "start from silence and discard the final state".
 *  - The library never allocates or frees device memory on the device entry points: the
 *    caller owns every buffer including `workspace` (query the size first).  Only the
 *    `_host` streaming entry points own (cached) device staging buffers.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream, which is what
 *    every launch of the reference uses: src/torchfx/_csrc/cuda/parallel_scan.cu:299-352).
 *  - Every function returns TFX_OK (0) or a negative TFX_E* code; tfx_last_error() gives
 *    the thread-local message.  A device entry point called on a box without a usable
 *    GPU returns TFX_ENODEVICE -- it never computes on the CPU instead.
 *  - y may alias x exactly (in place) on every entry point that says so.
 */
#ifndef TORCHFX_B200_H
#define TORCHFX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFX_VERSION 100 /* 0.1.0 */

/* ---- return codes ---------------------------------------------------------------- */
#define TFX_OK          0
#define TFX_EINVAL     -1 /* bad argument (shape, NULL pointer, K out of range, ...)      */
#define TFX_ENODEVICE  -2 /* no usable CUDA device / driver                              */
#define TFX_ECUDA      -3 /* a CUDA call or kernel launch failed (message has the string) */
#define TFX_EWORKSPACE -4 /* workspace too small                                         */
#define TFX_ENOMEM     -5 /* host or staging allocation failed                           */

/* ---- flags ------------------------------------------------------------------------ */
#define TFX_PREC_MASK   0x3u
#define TFX_PREC_AUTO   0x0u /* per-filter policy: f32 recurrence when a host-side probe of  */
                             /* this cascade bounds its round-off below TFX_AUTO_F32_BOUND,  */
                             /* else f64 recurrence (hard part 1)               */
#define TFX_PREC_F32    0x1u /* force float32 recurrence (f32 I/O only)                      */
#define TFX_PREC_F64    0x2u /* force float64 recurrence (what the reference computes in)    */
#define TFX_NO_TMA      0x8u  /* RESERVED, accepted and ignored: round 1 shipped a TMA-tiled and a  */
#define TFX_FORCE_TMA   0x10u /* packed-pair (FFMA2) cascade kernel behind these bits; both lost    */
#define TFX_PACKED      0x20u /* every A/B against the cp.async tile kernel on B200 and were removed */
                              /* (profiles/r1_experiments.md sections 3, 4, 12 keep the record)     */
#define TFX_NO_TILE     0x40u /* never take the channel-tile kernel (lanes = channels); use the    */
                              /* stream-per-lane kernel even for many channels (A/B, tests)       */
#define TFX_FORCE_TILE  0x80u /* filterbank: take the lanes = channels kernel (bank_stack) even when */
                              /* its grid would leave most SMs idle and the library would pick the   */
                              /* band-per-lane kernel (small C x T; A/B, tests)                      */
#define TFX_BANK_STRICT_ORDER 0x100u /* SUM banks of 9..32 children: add the children in child order, as the  */
                              /* reference's `out += f(x)` loop does (filter/__base.py:1019-1026), on   */
                              /* the band-per-lane kernel; default = per-warp partial sums (faster,     */
                              /* rounding-level difference).  Banks of <= 8 children always add in      */
                              /* child order                                                            */
#define TFX_NO_SPLIT    0x4u /* never split a channel in time (one sequential stream per     */
                             /* channel; exact for unstable filters; used by tests)          */

/* filterbank output mode */
#define TFX_BANK_STACK  0 /* y[b, c, t]  -- LogFilterBank (reference filterbank.py:183-185)   */
#define TFX_BANK_SUM    1 /* y[c, t] = sum_b -- ParallelFilterCombination (__base.py:1019-1026) */

/* FIR algorithm selection */
#define TFX_FIR_AUTO    0 /* direct up to 56 taps, and up to 96 taps below 16 M samples; else OLS     */
#define TFX_FIR_DIRECT  1 /* shared-memory tiled direct form                                  */
#define TFX_FIR_OLS     2 /* (partitioned) overlap-save block FFT in shared memory            */

/* ---- library ---------------------------------------------------------------------- */
int         tfx_version(void);
const char *tfx_last_error(void);
/* number of usable CUDA devices (0 on a CPU-only box; never fails) */
int         tfx_device_count(void);
/* total number of CUDA kernels this library has launched in this process (monotonic;   */
/* tests and bench.py read it to prove the device path ran)                              */
uint64_t    tfx_kernel_launches(void);

/* ==== SOS / biquad cascade ===========================================================
 * Replaces torchfx_ext.sos_forward and torchfx_ext.biquad_forward
 * (reference binding.cpp:30-66 -> cuda/biquad_forward.cu:7-92, cuda/parallel_scan.cu:117-364;
 * callers src/torchfx/_ops.py:57-176, src/torchfx/filter/iir.py:149-176).
 * One fused launch for all K sections (K >= 1; cascades longer than TFX_SOS_MAX_FUSED
 * sections run as ceil(K / TFX_SOS_MAX_FUSED) in-place passes).
 * In place (y == x, ldy == ldx) is supported.                                            */
#define TFX_SOS_MAX_FUSED 8
#define TFX_SOS_MAX_K     64

size_t tfx_sos_cascade_workspace_bytes(int64_t C, int64_t T, int K);

int tfx_sos_cascade_f32(const float *x, float *y, int64_t C, int64_t T,
                        int64_t ldx, int64_t ldy,
                        const double *sos_host, int K,
                        double *state_x, double *state_y,
                        uint32_t flags, void *workspace, size_t workspace_bytes,
                        void *stream);

int tfx_sos_cascade_f64(const double *x, double *y, int64_t C, int64_t T,
                        int64_t ldx, int64_t ldy,
                        const double *sos_host, int K,
                        double *state_x, double *state_y,
                        uint32_t flags, void *workspace, size_t workspace_bytes,
                        void *stream);

/* The planner's time segmentation for `lanes` streams of T samples whose filter needs `warm_needed`
 * warm-up samples (negative: never forgets -> no split) on a GPU that holds `capacity` streams in one
 * wave, `oversub` work items per resident warp wanted (1 = static partition).  Pure host arithmetic
 * (sos_plan.cpp); exported so that tests and integrators can see what a call will do without a
 * device (TFX_PLAN_DEBUG=1 prints the same numbers for live calls).  Outputs: segments per
 * channel, segment length (a multiple of 64 when split), warm-up length (multiple of 64).         */
void tfx_plan_segmentation(int64_t lanes, int64_t T, int64_t warm_needed, int64_t capacity, int oversub,
                           int64_t *segments, int64_t *segment_len, int64_t *warm);

/* What TFX_PREC_AUTO resolves to for this cascade: returns TFX_PREC_F32 or TFX_PREC_F64,
 * and (optionally) the probe's estimated f32 round-off relative to max|y|.               */
int tfx_sos_auto_precision(const double *sos_host, int K, double *probe_rel_err);
/* When the policy above answers TFX_PREC_F64: the sections (bit k = section k) that have to
 * run the float64 recurrence for the probe error to fall under the bound; the remaining
 * sections keep the float32 recurrence in the mixed-precision channel-tile kernel.  All K
 * bits set = no proper subset suffices (the whole cascade runs in float64).  0 for cascades
 * whose policy is TFX_PREC_F32.                                                          */
uint64_t tfx_sos_mixed_mask(const double *sos_host, int K, double *mixed_rel_err);

/* Host twins (HOST pointers, OpenMP over channels, f64 DF1 -- the arithmetic of the
 * reference's CPU kernel cpu/iir_cpu.cpp:64-159).  This is device dispatch, mirroring
 * binding.cpp:38,59 (`x.is_cuda()`), not a fallback: CUDA tensors never reach it.        */
int tfx_sos_cascade_cpu_f32(const float *x, float *y, int64_t C, int64_t T,
                            int64_t ldx, int64_t ldy, const double *sos_host, int K,
                            double *state_x, double *state_y);
int tfx_sos_cascade_cpu_f64(const double *x, double *y, int64_t C, int64_t T,
                            int64_t ldx, int64_t ldy, const double *sos_host, int K,
                            double *state_x, double *state_y);

/* Host-buffer streaming driver (reference caller
 * src/torchfx/realtime/stream.py:279-347): x_host / y_host are HOST buffers (pinned for
 * full speed), processed on `device` in time chunks of `chunk_T` samples (0 = default)
 * with H2D, kernel and D2H overlapped on three streams and the DF1 state carried from
 * chunk to chunk on the device.  state_x / state_y are HOST [K, C, 2] or NULL.
 * This is the call bench.py times for its `e2e` number.                                  */
int tfx_sos_cascade_host_f32(const float *x_host, float *y_host, int64_t C, int64_t T,
                             int64_t ldx, int64_t ldy, const double *sos_host, int K,
                             double *state_x_host, double *state_y_host,
                             uint32_t flags, int64_t chunk_T, int device);

/* ==== parallel filterbank ============================================================
 * N cascades of Kb sections each applied to the SAME input in one launch.
 * Replaces the Python loops of LogFilterBank.forward (reference filterbank.py:183-185,
 * mode STACK: y is [N, C, T], band stride `ldb` elements) and
 * ParallelFilterCombination.forward (__base.py:1019-1026, mode SUM: y is [C, T]).
 * sos_host is [N, Kb, 6]; state_x / state_y are [N, Kb, C, 2] device doubles or NULL.    */
#define TFX_BANK_MAX_LANES 64 /* N * Kb per launch; more bands run as several launches   */

size_t tfx_filterbank_workspace_bytes(int64_t C, int64_t T, int N, int Kb);

int tfx_filterbank_f32(const float *x, float *y, int64_t C, int64_t T,
                       int64_t ldx, int64_t ldy, int64_t ldb,
                       const double *sos_host, int N, int Kb, int mode,
                       double *state_x, double *state_y,
                       uint32_t flags, void *workspace, size_t workspace_bytes,
                       void *stream);

int tfx_filterbank_f64(const double *x, double *y, int64_t C, int64_t T,
                       int64_t ldx, int64_t ldy, int64_t ldb,
                       const double *sos_host, int N, int Kb, int mode,
                       double *state_x, double *state_y,
                       uint32_t flags, void *workspace, size_t workspace_bytes,
                       void *stream);

/* ==== FIR ============================================================================
 * Causal FIR with zero history, output length T: y[c,n] = sum_j taps[j] * x[c,n-j].
 * Replaces FIR.forward -> fft_conv1d / F.conv1d (reference filter/fir.py:526-579,
 * filter/_fftconv.py:107-141; torch.fft / cuFFT are not used).  `taps` is a DEVICE
 * pointer to the K impulse-response samples b[0..K) in natural order (the reference
 * stores them flipped, fir.py:516-518; the Python layer un-flips).  Not in place.        */
size_t tfx_fir_workspace_bytes(int64_t C, int64_t T, int64_t K, int algo);

int tfx_fir_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                const float *taps, int64_t K, int algo,
                void *workspace, size_t workspace_bytes, void *stream);

/* Plan for the overlap-save algorithm: everything that depends on the impulse response only
 * (FFT twiddle tables and the spectra of the K-tap response's partitions) computed ONCE into
 * a caller-owned, 16-byte aligned DEVICE buffer of tfx_fir_plan_bytes(K) bytes, so that
 * chunked callers (a FIR module called per audio block) do not transform the taps on every
 * call.  The reference keeps its flipped kernel as a module buffer (filter/fir.py:516-518)
 * and re-runs rfft(kernel) inside every fft_conv1d call (filter/_fftconv.py:124).
 * tfx_fir_f32_planned == tfx_fir_f32 with algo TFX_FIR_OLS and the same workspace
 * (tfx_fir_workspace_bytes(C, T, K, TFX_FIR_OLS)); results are bit-identical.             */
size_t tfx_fir_plan_bytes(int64_t K);
int tfx_fir_plan_init(const float *taps, int64_t K, void *plan, size_t plan_bytes, void *stream);
int tfx_fir_f32_planned(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        const void *plan, int64_t K,
                        void *workspace, size_t workspace_bytes, void *stream);

/* float64 signals on the device: the reference evaluates the sum in the input dtype
 * (filter/fir.py:529-531).  Direct form in float64 for every K (no workspace).            */
int tfx_fir_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                const double *taps, int64_t K, void *stream);

int tfx_fir_cpu_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                    const float *taps_host, int64_t K);
int tfx_fir_cpu_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                    const double *taps_host, int64_t K);

/* ==== delay line =====================================================================
 * y[n] = x[n] + mix*decay*x[n-delay] (n >= delay), y[n] = x[n] (n < delay).
 * Replaces torchfx_ext.delay_line_forward (reference binding.cpp:68-81,
 * cuda/delay_forward.cu:15-123, cpu/delay_cpu.cpp:17-85).  Not in place.                 */
int tfx_delay_line_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                       int64_t delay, double decay, double mix, void *stream);
int tfx_delay_line_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                       int64_t delay, double decay, double mix, void *stream);
int tfx_delay_line_cpu_f32(const float *x, float *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                           int64_t delay, double decay, double mix);
int tfx_delay_line_cpu_f64(const double *x, double *y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                           int64_t delay, double decay, double mix);

#ifdef __cplusplus
}
#endif
#endif /* TORCHFX_B200_H */
