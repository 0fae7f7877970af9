// Microbenchmark (developer tool): issue throughput of FFMA vs FFMA2 (packed f32x2, sm_100) alone and
// interleaved with integer ALU work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_issue fp32x2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int n) {
    float x[8];
    float2 y[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = make_float2(x[i], x[i] + 1.f); u[i] = threadIdx.x * 7 + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = __fmaf_rn(x[i], a, b);                       // 8 FFMA
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);                    // 8 FFMA2
            if (MODE == 2) { x[i] = __fmaf_rn(x[i], a, b); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }   // FFMA + 2 ALU (LOP3/SHF + IADD)
            if (MODE == 3) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }
            if (MODE == 4) { x[i] = __fmaf_rn(x[i], a, b); y[i].x = __fmaf_rn(y[i].x, a, b); }            // 16 FFMA
            if (MODE == 5) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u;             // ALU only
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i].x + y[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int inst_per_iter) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, ITERS);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, ITERS);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = 148.0 * 8 * 8 /*warps*/ * (double)ITERS * inst_per_iter;
    printf("%-28s %8.3f ms  %7.1f G warp-inst/s  (%.2f per clk per SMSP @1.965GHz)\n", name, ms, winst / ms / 1e6,
           winst / (ms * 1e-3) / (592 * 1.965e9));
    cudaFree(out);
}

int main() {
    run<0>("8xFFMA", 8);
    run<1>("8xFFMA2", 8);
    run<2>("8x(FFMA+3ALU)", 32);
    run<3>("8x(FFMA2+3ALU)", 32);
    run<4>("16xFFMA", 16);
    run<5>("8x(3ALU)", 24);
    return 0;
}
