// Microbenchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 issue throughput on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_fp32.cu -o tools/ubench_fp32
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = fmaf(x[i].y, aa.y, bb.y); }     // 2 FFMA
            if (MODE == 1) { x[i] = __ffma2_rn(x[i], aa, bb); }                                          // 1 FFMA2
            if (MODE == 2) { x[i].x = x[i].x + bb.x; x[i].y = x[i].y + bb.y; }                           // 2 FADD
            if (MODE == 3) { x[i] = __fadd2_rn(x[i], bb); }                                              // 1 FADD2
            if (MODE == 4) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = x[i].y + bb.y; }                // FFMA + FADD
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out) {
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)grid * 256 * iters * 16;       // 16 scalar fp32 ops per thread-iteration
    printf("%-14s %8.3f ms  %7.2f T scalar-op/s  (%.2f ops/clk/SM at 1.965 GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out); run<4>("FFMA+FADD", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
