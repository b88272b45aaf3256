// Are the FP64 FMA pipe (DFMA) and the FP64 tensor pipe (DMMA.8x8x4) of sm_100a independent?  Times three kernels with the
// same number of warps: DFMA only, DMMA only, and both interleaved in every warp.  If the mixed kernel takes ~max(t1, t2)
// the pipes overlap and a GEMM could feed both; if it takes ~t1 + t2 they share the datapath.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mb_fp64_pipes scripts/mb_fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DFMA, 1: DMMA, 2: both
__global__ void __launch_bounds__(256) pipes(double* out, int iters, double x) {
    double f[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = x + i; c[i][0] = x; c[i][1] = x; }
    const double a = x * 1.0000001, b = x * 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 1) {
                // 16 DFMA per lane = 32 x 16 x 2 = 1024 flops per warp, the same as two DMMA.8x8x4
#pragma unroll
                for (int r = 0; r < 2; ++r) f[i] = fma(f[i], a, b);
            }
            if (MODE != 0) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> double run(double* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipes<MODE><<<148 * 8, 256>>>(out, iters, 1.0);
    cudaEventRecord(e0);
    pipes<MODE><<<148 * 8, 256>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    const int iters = 20000;
    const double warps = 148.0 * 8 * 8;
    const double t0 = run<0>(out, iters), t1 = run<1>(out, iters), t2 = run<2>(out, iters);
    const double dfma_flops = warps * iters * 8.0 * 2 * 32 * 2;      // per warp and iteration: 8 x 2 DFMA x 32 lanes x 2 flops
    const double dmma_flops = warps * iters * 8.0 * 512;             // 8 DMMA.8x8x4 of 512 flops
    printf("DFMA only : %8.3f ms  %6.2f TFLOP/s\n", t0, dfma_flops / t0 / 1e9);
    printf("DMMA only : %8.3f ms  %6.2f TFLOP/s\n", t1, dmma_flops / t1 / 1e9);
    printf("both      : %8.3f ms  %6.2f TFLOP/s  (sum of the two alone: %.3f ms, max: %.3f ms)\n", t2, (dfma_flops + dmma_flops) / t2 / 1e9,
           t0 + t1, t0 > t1 ? t0 : t1);
    return 0;
}
