// Shared helpers of the tnsp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>

#include "tnsp_b200.h"

namespace tnsp {

extern std::atomic<int64_t> g_launches;
void set_error(const std::string& what);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        return 1;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

constexpr int kSMs = 148;  // B200

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` must hold >= 32 doubles; result broadcast to all threads.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (warp == 0) {
        r = warp_sum(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ double block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (warp == 0) {
        r = warp_max(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    return red[0];
}

}  // namespace tnsp
