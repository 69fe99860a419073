// Developer probe: achievable HBM WRITE bandwidth for the store pattern of a tile kernel --
// every warp round-robins over `rows` rows (row stride = row length), writing `piece` bytes of
// each row per round with 16-byte streaming stores -- as a function of piece size, rows per warp
// and footprint.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_pattern write_pattern.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void st16(void *p, float4 v, int cs) {
    if (cs)
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else
        *reinterpret_cast<float4 *>(p) = v;
}

// rows_per_warp rows, piece_bytes per row per round (multiple of 256 or 512), row_bytes per row
__global__ void __launch_bounds__(256, 2) wr(unsigned char *y, int64_t row_bytes, int rows, int piece_bytes, int cs, int delay) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    unsigned char *base = y + w * rows * row_bytes;
    const float4 v = make_float4(1.f, 2.f, 3.f, float(lane));
    const int64_t rounds = row_bytes / piece_bytes;
    for (int64_t i = 0; i < rounds; ++i) {
        for (int r = 0; r < rows; ++r) {
            unsigned char *p = base + r * row_bytes + i * piece_bytes;
            for (int o = lane * 16; o < piece_bytes; o += 512) st16(p + o, v, cs);
        }
        if (delay) __nanosleep(delay);
    }
}
// piece = 256: two rows per instruction (half-warp each), like the kernels
__global__ void __launch_bounds__(256, 2) wr256(unsigned char *y, int64_t row_bytes, int rows, int cs) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    unsigned char *base = y + w * rows * row_bytes + (lane >> 4) * row_bytes + (lane & 15) * 16;
    const float4 v = make_float4(1.f, 2.f, 3.f, float(lane));
    const int64_t rounds = row_bytes / 256;
    for (int64_t i = 0; i < rounds; ++i)
        for (int r = 0; r < rows; r += 2) st16(base + r * row_bytes + i * 256, v, cs);
}

int main(int argc, char **argv) {
    const int ctas = 296, warps = ctas * 8;
    unsigned char *y;
    const size_t cap = size_t(100) << 30;
    if (cudaMalloc(&y, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, int rows, int piece, double gb, int cs) {
        int64_t row_bytes = int64_t(gb * 1e9 / (double(warps) * rows));
        row_bytes = row_bytes / 4096 * 4096;
        const double bytes = double(row_bytes) * rows * warps;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (piece == 256) wr256<<<ctas, 256>>>(y, row_bytes, rows, cs);
            else wr<<<ctas, 256>>>(y, row_bytes, rows, piece, cs, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%-28s rows/warp %3d piece %5d B footprint %5.1f GB cs=%d : %7.3f ms  %6.0f GB/s\n", name, rows, piece, bytes / 1e9, cs, best, bytes / best / 1e6);
        fflush(stdout);
    };
    for (double gb : {94.0, 24.0}) {
        run("tile pattern", 32, 256, gb, 1);
        run("tile pattern (plain st)", 32, 256, gb, 0);
        run("wider pieces", 32, 512, gb, 1);
        run("wider pieces", 32, 1024, gb, 1);
        run("wider pieces", 32, 4096, gb, 1);
        run("fewer rows", 8, 256, gb, 1);
        run("fewer rows", 8, 512, gb, 1);
        run("fewer rows", 4, 512, gb, 1);
        run("one row per warp", 1, 512, gb, 1);
        run("one row per warp", 1, 4096, gb, 1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
