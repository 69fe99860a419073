// host_stream.cu -- host-buffer streaming driver for the SOS cascade.
//
// The caller hands HOST buffers (the reference's users hold audio in host memory:
// Wave.from_file -> CPU tensor, src/torchfx/wave.py:406-470; long files are processed in
// chunks with state carry-over, src/torchfx/realtime/stream.py:279-347).  The signal is cut
// into time chunks; chunk i is copied H2D, filtered IN PLACE on the device by the fused
// cascade kernel, and copied back D2H, on three streams so that the copy-in of chunk i+1,
// the kernel of chunk i and the copy-out of chunk i-1 overlap (PCIe is full duplex).  The
// DF1 state lives on the device between chunks -- the reference's own chunk contract
// (filter/iir.py:135-144), so the result equals one unbroken call.
//
// This is the only entry point that owns device memory (a per-process staging cache).
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tfx {
namespace {

constexpr int kBufs = 3;

struct Staging {
    int device = -1;
    float *buf[kBufs] = {nullptr, nullptr, nullptr};
    size_t buf_elems = 0;
    double *state = nullptr;  // [2][K, C, 2]
    size_t state_elems = 0;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t e_in[kBufs], e_k[kBufs], e_out[kBufs];
    bool init = false;
    std::mutex mu;  // one streaming call at a time PER DEVICE (the staging buffers and streams are per device)
};

std::mutex g_mu;
std::vector<Staging *> g_staging;

Staging *staging_for(int device) {
    for (Staging *s : g_staging)
        if (s->device == device) return s;
    Staging *s = new Staging();
    s->device = device;
    g_staging.push_back(s);
    return s;
}

int ensure(Staging &s, size_t buf_elems, size_t state_elems, size_t ws_bytes) {
    if (!s.init) {
        TFX_CUDA_TRY(cudaStreamCreateWithFlags(&s.s_in, cudaStreamNonBlocking));
        TFX_CUDA_TRY(cudaStreamCreateWithFlags(&s.s_k, cudaStreamNonBlocking));
        TFX_CUDA_TRY(cudaStreamCreateWithFlags(&s.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kBufs; ++i) {
            TFX_CUDA_TRY(cudaEventCreateWithFlags(&s.e_in[i], cudaEventDisableTiming));
            TFX_CUDA_TRY(cudaEventCreateWithFlags(&s.e_k[i], cudaEventDisableTiming));
            TFX_CUDA_TRY(cudaEventCreateWithFlags(&s.e_out[i], cudaEventDisableTiming));
        }
        s.init = true;
    }
    if (buf_elems > s.buf_elems) {
        for (int i = 0; i < kBufs; ++i) {
            if (s.buf[i]) TFX_CUDA_TRY(cudaFree(s.buf[i]));
            s.buf[i] = nullptr;
        }
        s.buf_elems = 0;
        for (int i = 0; i < kBufs; ++i) TFX_CUDA_TRY(cudaMalloc(&s.buf[i], buf_elems * sizeof(float)));
        s.buf_elems = buf_elems;
    }
    if (state_elems > s.state_elems) {
        if (s.state) TFX_CUDA_TRY(cudaFree(s.state));
        s.state = nullptr;
        s.state_elems = 0;
        TFX_CUDA_TRY(cudaMalloc(&s.state, state_elems * sizeof(double)));
        s.state_elems = state_elems;
    }
    if (ws_bytes > s.ws_bytes) {
        if (s.ws) TFX_CUDA_TRY(cudaFree(s.ws));
        s.ws = nullptr;
        s.ws_bytes = 0;
        TFX_CUDA_TRY(cudaMalloc(&s.ws, ws_bytes));
        s.ws_bytes = ws_bytes;
    }
    return TFX_OK;
}

}  // namespace
}  // namespace tfx

extern "C" int tfx_sos_cascade_host_f32(const float *x_host, float *y_host, int64_t C, int64_t T, int64_t ldx,
                                        int64_t ldy, const double *sos_host, int K, double *state_x_host,
                                        double *state_y_host, uint32_t flags, int64_t chunk_T, int device) {
    using namespace tfx;
    TFX_REQUIRE(C >= 0 && T >= 0, "sos cascade (host): negative shape");
    TFX_REQUIRE(K >= 1 && K <= TFX_SOS_MAX_K && sos_host != nullptr, "sos cascade (host): bad K / sos");
    TFX_REQUIRE((state_x_host == nullptr) == (state_y_host == nullptr), "sos cascade (host): state_x and state_y must both be given or both NULL");
    if (C == 0 || T == 0) return TFX_OK;
    TFX_REQUIRE(x_host != nullptr && y_host != nullptr, "sos cascade (host): NULL signal pointer");
    TFX_REQUIRE(ldx >= T && ldy >= T, "sos cascade (host): row stride smaller than T");
    int rc = require_device();
    if (rc != TFX_OK) return rc;

    int prev = 0;
    TFX_CUDA_TRY(cudaGetDevice(&prev));
    TFX_CUDA_TRY(cudaSetDevice(device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev};

    if (chunk_T <= 0) {
        // ~256 MB per staging buffer: long enough that every time segment dwarfs its warm-up.
        chunk_T = std::max<int64_t>(int64_t(64) << 20, 1) / std::max<int64_t>(C, 1);
        chunk_T = std::max<int64_t>(chunk_T, 16384);
    }
    chunk_T = std::min(chunk_T, T);
    chunk_T = (chunk_T + 3) / 4 * 4;

    Staging *sp = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);  // the registry only; calls on different devices run concurrently
        sp = staging_for(device);
    }
    Staging &s = *sp;
    std::lock_guard<std::mutex> lk(s.mu);
    const size_t state_elems = static_cast<size_t>(K) * C * 2;
    const size_t ws_bytes = tfx_sos_cascade_workspace_bytes(C, chunk_T, K);
    rc = ensure(s, static_cast<size_t>(C) * chunk_T, 2 * state_elems, ws_bytes);
    if (rc != TFX_OK) return rc;

    double *d_sx = s.state, *d_sy = s.state + state_elems;
    if (state_x_host != nullptr) {
        TFX_CUDA_TRY(cudaMemcpyAsync(d_sx, state_x_host, state_elems * sizeof(double), cudaMemcpyHostToDevice, s.s_k));
        TFX_CUDA_TRY(cudaMemcpyAsync(d_sy, state_y_host, state_elems * sizeof(double), cudaMemcpyHostToDevice, s.s_k));
    } else {
        TFX_CUDA_TRY(cudaMemsetAsync(s.state, 0, 2 * state_elems * sizeof(double), s.s_k));
    }

    const int64_t nchunks = (T + chunk_T - 1) / chunk_T;
    for (int64_t i = 0; i < nchunks; ++i) {
        const int b = static_cast<int>(i % kBufs);
        const int64_t t0 = i * chunk_T;
        const int64_t len = std::min(chunk_T, T - t0);
        // copy-in may reuse buffer b only after the copy-out of chunk i - kBufs
        if (i >= kBufs) TFX_CUDA_TRY(cudaStreamWaitEvent(s.s_in, s.e_out[b], 0));
        TFX_CUDA_TRY(cudaMemcpy2DAsync(s.buf[b], chunk_T * sizeof(float), x_host + t0, ldx * sizeof(float),
                                       len * sizeof(float), C, cudaMemcpyHostToDevice, s.s_in));
        TFX_CUDA_TRY(cudaEventRecord(s.e_in[b], s.s_in));
        TFX_CUDA_TRY(cudaStreamWaitEvent(s.s_k, s.e_in[b], 0));
        rc = tfx_sos_cascade_f32(s.buf[b], s.buf[b], C, len, chunk_T, chunk_T, sos_host, K, d_sx, d_sy, flags, s.ws,
                                 s.ws_bytes, s.s_k);
        if (rc != TFX_OK) {
            cudaDeviceSynchronize();
            return rc;
        }
        TFX_CUDA_TRY(cudaEventRecord(s.e_k[b], s.s_k));
        TFX_CUDA_TRY(cudaStreamWaitEvent(s.s_out, s.e_k[b], 0));
        TFX_CUDA_TRY(cudaMemcpy2DAsync(y_host + t0, ldy * sizeof(float), s.buf[b], chunk_T * sizeof(float),
                                       len * sizeof(float), C, cudaMemcpyDeviceToHost, s.s_out));
        TFX_CUDA_TRY(cudaEventRecord(s.e_out[b], s.s_out));
    }
    if (state_x_host != nullptr) {
        TFX_CUDA_TRY(cudaMemcpyAsync(state_x_host, d_sx, state_elems * sizeof(double), cudaMemcpyDeviceToHost, s.s_k));
        TFX_CUDA_TRY(cudaMemcpyAsync(state_y_host, d_sy, state_elems * sizeof(double), cudaMemcpyDeviceToHost, s.s_k));
    }
    TFX_CUDA_TRY(cudaStreamSynchronize(s.s_in));
    TFX_CUDA_TRY(cudaStreamSynchronize(s.s_k));
    TFX_CUDA_TRY(cudaStreamSynchronize(s.s_out));
    return TFX_OK;
}
