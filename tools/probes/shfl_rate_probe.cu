// Probe: warp-shuffle issue rate per SM (cycles per SHFL warp-instruction) for 4, 8 and 16 resident warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o shfl_rate_probe.bin shfl_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(long long *out, int iters, float *sink)
{
    const int lane = threadIdx.x & 31;
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = threadIdx.x * 0.5f + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(0xFFFFFFFFu, v[j], (lane + 1) & 31);     // 16 independent shuffles
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += v[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main()
{
    long long *d, h;
    float *sink;
    cudaMalloc(&d, 8 * 148);
    cudaMalloc(&sink, 4 * 1024 * 148);
    for (int warps : {1, 4, 8, 16, 32}) {
        const int iters = 200;
        k<<<1, 32 * warps>>>(d, iters, sink);
        cudaDeviceSynchronize();
        k<<<1, 32 * warps>>>(d, iters, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d: %lld cycles for %d shuffles per warp -> %.2f cycles per warp-shuffle per SM\n", warps, h, iters * 16,
               (double)h / (iters * 16.0 * warps));
    }
    return 0;
}
