// Probe (round 2): what bounds a tcgen05.mma with a NARROW N, and can the A operand escape the shared-memory fetch?
//   (1) A from shared memory, M = 128 (the product trunk's shape): cycles per UMMA at N = 32 / 64      [baseline]
//   (2) A from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): same shapes — no A fetch from shared memory
//   (3) M = 64 from shared memory: half the A bytes per instruction
// Four issuing threads, one accumulator each (the aggregate rate of the tensor pipe, as in umma_rate_probe.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_atmem_probe umma_atmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tacult_b200/csrc/tcgen05.cuh"

using namespace tc;

__device__ __forceinline__ void umma_bf16_atmem(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ uint32_t idesc_mn(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode 0: A smem M=128; 1: A tmem M=128; 2: A smem M=64
__global__ void k(long long *out, int mode, int n_cols, int iters)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars4[4];
    __shared__ long long t_end[4];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) mbar_init(bars4 + w, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    const int w = threadIdx.x >> 5;
    constexpr int SW = 64;
    __syncthreads();
    const long long t0 = clock64();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        const uint32_t idesc = idesc_mn(mode == 2 ? 64 : 128, n_cols);
        // accumulators: 4 x n_cols columns from 0; A operands in TMEM (mode 1): columns [448, 512), 8 columns per K=16 slice
        const uint32_t d = base + (uint32_t)(w * (n_cols <= 96 ? 96 : n_cols));
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = a0 + (uint32_t)(((i + 2 * w) & 7) * 128 * SW) + (uint32_t)(((i >> 3) & 1) * 32);
            const uint64_t bd = make_smem_desc<SW>(b0 + (uint32_t)(((i + w) & 3) * 256 * SW));
            if (mode == 1) umma_bf16_atmem(d, base + 448u + (uint32_t)((i & 7) * 8), bd, idesc, 1u);
            else umma_bf16(d, make_smem_desc<SW>(a), bd, idesc, 1u);
        }
        umma_commit(bars4 + w);
        mbar_wait(bars4 + w, 0);
        t_end[w] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = max(max(t_end[0], t_end[1]), max(t_end[2], t_end[3]));
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

int main()
{
    long long *d, h = 0;
    cudaMalloc(&d, 64);
    const int iters = 2000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char *names[3] = {"A from shared memory, M = 128", "A from TENSOR memory, M = 128", "A from shared memory, M =  64"};
    for (int mode = 0; mode < 3; ++mode)
        for (int n : {32, 64, 96}) {
            if (4 * (n <= 96 ? 96 : n) > 448) continue;
            cudaMemset(d, 0, 64);
            k<<<1, 128, 196 * 1024>>>(d, mode, n, iters);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("%s, N = %3d, K = 16: %s  %.1f cycles per UMMA (four issuers, aggregate)  -> %.0f MAC/clk\n", names[mode], n,
                   cudaGetErrorString(e), (double)h / (4.0 * iters), (mode == 2 ? 64.0 : 128.0) * n * 16 / ((double)h / (4.0 * iters)));
        }
    return 0;
}
