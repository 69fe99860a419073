// Developer probe: issue rates of DFMA, F2F.F64.F32, F2F.F32.F64 and their mix on sm_100a, normalised to FFMA
// (128 lanes/clk/SM), plus an integer-ALU float->double widening.  Answers: do the conversions share the FP64 pipe?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f64_rates f64_rates.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__device__ __forceinline__ double widen_alu(float f) {  // exact for normal floats and zero; denormals flush to zero
    const unsigned u = __float_as_uint(f);
    const unsigned e = (u >> 23) & 0xffu;
    unsigned hi = (u & 0x80000000u) | (((e + 896u) << 20) | ((u >> 3) & 0xfffffu));
    unsigned lo = u << 29;
    if (e == 0u) { hi = u & 0x80000000u; lo = 0u; }
    return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float seed) {
    float f[ILP];
    double d[ILP];
    for (int i = 0; i < ILP; ++i) { f[i] = seed + i + threadIdx.x; d[i] = f[i] * 0.5; }
    const double a = 0.999999, b = 1e-9;
    const float af = 0.999999f, bf = 1e-9f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) f[i] = fmaf(f[i], af, bf);                       // FFMA
            if (MODE == 1) d[i] = fma(d[i], a, b);                          // DFMA
            if (MODE == 2) { d[i] = static_cast<double>(f[i]); f[i] = __double_as_longlong(d[i]) & 1 ? f[i] + 1.f : f[i]; }  // F2F.F64.F32 (+ light ALU)
            if (MODE == 3) { f[i] = static_cast<float>(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]), __float_as_int(f[i])); }  // F2F.F32.F64
            if (MODE == 4) { d[i] = fma(d[i], a, static_cast<double>(f[i])); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); f[i] = static_cast<float>(d[i]); }  // 5 DFMA + 2 F2F
            if (MODE == 5) { d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); }  // 5 DFMA
            if (MODE == 6) { d[i] = fma(d[i], a, widen_alu(f[i])); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); f[i] = static_cast<float>(d[i]); }  // 5 DFMA + ALU widen + 1 F2F
            if (MODE == 7) { d[i] = widen_alu(f[i]); f[i] = __double2hiint(d[i]) & 1 ? f[i] + 1.f : f[i]; }  // ALU widen only
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += f[i] + static_cast<float>(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(float *out, const char *name, double ops_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double units = double(blocks) * 256 * ITERS * ILP;  // "samples"
    printf("%-44s %8.3f ms  %8.1f G units/s  (%5.1f G %s/s)\n", name, ms, units / ms / 1e6, units * ops_per_iter / ms / 1e6, "ops");
    return units / ms / 1e6;
}

int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const double ffma = run<0>(out, "FFMA", 1);
    const double per_clk = ffma / 128.0;  // G (SM*clk)/s: FFMA issues 128 lanes/clk/SM
    printf("=> %.1f G SM-clocks/s (148 SMs at %.0f MHz)\n", per_clk, per_clk / 148 * 1e3);
    auto lanes = [&](double g, double ops) { return g * ops / per_clk; };
    double g;
    g = run<1>(out, "DFMA", 1);                                   printf("   DFMA lanes/clk/SM: %.1f\n", lanes(g, 1));
    g = run<2>(out, "F2F.F64.F32 (+1 ALU op)", 1);                printf("   widen lanes/clk/SM: %.1f\n", lanes(g, 1));
    g = run<3>(out, "F2F.F32.F64 (+1 ALU op)", 1);                printf("   narrow lanes/clk/SM: %.1f\n", lanes(g, 1));
    g = run<5>(out, "5 DFMA per unit", 5);                        printf("   SM-clk per warp-unit: %.2f\n", 32.0 / (g / per_clk));
    g = run<4>(out, "5 DFMA + widen F2F + narrow F2F per unit", 7); printf("   SM-clk per warp-unit: %.2f\n", 32.0 / (g / per_clk));
    g = run<6>(out, "5 DFMA + ALU widen + narrow F2F per unit", 7); printf("   SM-clk per warp-unit: %.2f\n", 32.0 / (g / per_clk));
    g = run<7>(out, "ALU widen only (+1 ALU op)", 1);             printf("   ALU widen lanes/clk/SM: %.1f\n", lanes(g, 1));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
