// Probe: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N and of the shared-memory layout of the
// operands (rows of 32 / 64 / 128 bytes with the matching swizzle).  Four accumulators are used round-robin and the A
// start address walks over 8 positions so that consecutive UMMAs never share an operand.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe.bin umma_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tacult_b200/csrc/tcgen05.cuh"

using namespace tc;

template <int SW>
__global__ void k(long long *out, int n_cols, int iters)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar, bars4[4];
    __shared__ long long t_end[4];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int w = 0; w < 4; ++w) mbar_init(bars4 + w, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        const uint32_t idesc = make_instr_desc(n_cols);
        uint32_t ph = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t a = a0 + (uint32_t)((i & 7) * 128 * SW) + (uint32_t)(((i >> 3) & (SW / 32 - 1)) * 32);
                umma_bf16(base + (uint32_t)(n_cols > 128 ? (i & 1) * 256 : (i & 3) * 128), make_smem_desc<SW>(a), make_smem_desc<SW>(b0), idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, ph);
            ph ^= 1;
            out[rep] = clock64() - t0;
        }
    }
    // ---- four issuing threads (one per warp), one accumulator each: aggregate rate of the tensor pipe ----
    __syncthreads();
    if (n_cols <= 128) {
        const int w = threadIdx.x >> 5;
        __syncthreads();
        const long long t0 = clock64();
        if ((threadIdx.x & 31) == 0) {
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
            const uint32_t idesc = make_instr_desc(n_cols);
            for (int i = 0; i < iters; ++i) {
                const uint32_t a = a0 + (uint32_t)(((i + 2 * w) & 7) * 128 * SW) + (uint32_t)(((i >> 3) & (SW / 32 - 1)) * 32);
                umma_bf16(base + (uint32_t)(w * 128), make_smem_desc<SW>(a), make_smem_desc<SW>(b0), idesc, 1u);
            }
            umma_commit(bars4 + w);
            mbar_wait(bars4 + w, 0);
            t_end[w] = clock64() - t0;
        }
        __syncthreads();
        if (threadIdx.x == 0) out[3] = max(max(t_end[0], t_end[1]), max(t_end[2], t_end[3]));
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

template <int SW>
void run(long long *d, int n_cols)
{
    long long h[4] = {0, 0, 0, 0};
    const int iters = 2000;
    cudaFuncSetAttribute(k<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    k<SW><<<1, 128, 165 * 1024>>>(d, n_cols, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("rows of %3d B, N = %3d: %s  one issuer %.1f cycles per UMMA; four issuers %.1f cycles per UMMA (aggregate)\n", SW, n_cols,
           cudaGetErrorString(e), (double)h[2] / iters, n_cols <= 128 ? (double)h[3] / (4.0 * iters) : 0.0);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    for (int n : {16, 32, 64, 96, 128, 256}) {
        run<32>(d, n);
        run<64>(d, n);
        run<128>(d, n);
    }
    return 0;
}
