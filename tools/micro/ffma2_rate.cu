// Developer microbenchmark: issue rate of scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2 / add.f32x2)
// on sm_100a.  Decides whether FP-bound kernels (FIR MAC, FFT butterflies) should use the packed forms.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long p2;
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
    float a[16]; p2 q[16];
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; q[i] = (p2)__float_as_uint(a[i]) | ((p2)__float_as_uint(a[i] + 1.f) << 32); }
    const float b = s, c = 0.5f * s;
    const p2 b2 = (p2)__float_as_uint(b) | ((p2)__float_as_uint(b) << 32), c2 = (p2)__float_as_uint(c) | ((p2)__float_as_uint(c) << 32);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);
            if (MODE == 1) a[i] = a[i] + c;
            if (MODE == 2) q[i] = fma2(q[i], b2, c2);
            if (MODE == 3) q[i] = add2(q[i], c2);
            if (MODE == 4) a[i] = fmaf(a[i], b, a[(i + 1) & 15]);   // three distinct register operands
            if (MODE == 5) q[i] = fma2(q[i], b2, q[(i + 1) & 15]);
        }
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i] + __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)(q[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name, int flops_per_op) {
    int sm = 148; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sm * 8 * 256);
    const int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sm * 8, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(e0); k<MODE><<<sm * 8, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = double(sm) * 8 * 256 * 16.0 * iters;  // lane-ops
    printf("%-28s %8.3f ms  %7.2f T lane-ops/s  %7.2f TFLOP/s  (%.1f lane-ops/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e9, ops * flops_per_op / ms / 1e9,
           ops / (ms * 1e-3) / sm / 1.9e9);
    cudaFree(out);
}
int main() {
    run<0>("FFMA  a=a*b+c (2 uniform)", 2); run<4>("FFMA  3 distinct regs", 2); run<1>("FADD", 1);
    run<2>("FFMA2 a=a*b+c", 4); run<5>("FFMA2 3 distinct regs", 4); run<3>("FADD2", 2);
    return 0;
}
