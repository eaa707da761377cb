// Probe: tcgen05.mma.cta_group::2 (a CTA pair, M = 256 = 128 rows of each CTA) with N = 96, K = 16 bf16 -
// (1) semantics: which half of the B tile each CTA of the pair supplies, where the results land;
// (2) rate: cycles per pair-UMMA from one issuing thread of the leader (the single-CTA instruction costs
//     (4096 + 32 N) / 128 = 56 cycles of shared-memory operand fetch at N = 96; the pair should fetch A 4 KB + half of B).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_pair_probe.bin umma_pair_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../tacult_b200/csrc/tcgen05.cuh"

using namespace tc;

constexpr int N = 96;

__device__ __forceinline__ uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }
__device__ __forceinline__ uint64_t desc32(uint32_t saddr)
{
    constexpr uint32_t hi = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);
    return ((uint64_t)hi << 32) | (uint64_t)(saddr >> 4);
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__host__ __device__ inline float a_val(int cta, int r, int k) { return (float)(((cta * 128 + r) % 7) - 3 + (k % 3)); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n % 5) - 2 + (k % 4)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(float *out, long long *cycles, int iters, int b_mode)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *A = smem;                   // 8 tiles of [128 x 16] bf16, rows of 32 bytes, 32-byte swizzle (tile 0 holds the test data)
    uint8_t *B = smem + 8 * 4096;        // this CTA's part of the [96 x 16] B tile
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (8 * 4096 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 128 * 16; idx += blockDim.x) {
        const int r = idx >> 4, kk = idx & 15;
        for (int t = 0; t < 8; ++t)
            *reinterpret_cast<__nv_bfloat16 *>(A + t * 4096 + swz32((uint32_t)(r * 32 + kk * 2))) = __float2bfloat16_rn(a_val((int)rank, r, kk));
    }
    // b_mode 0: this CTA holds rows [48 rank, 48 rank + 48) of B in its rows 0..47;  1: both CTAs hold all 96 rows
    for (int idx = threadIdx.x; idx < N * 16; idx += blockDim.x) {
        const int n = idx >> 4, kk = idx & 15;
        if (b_mode == 0) {
            if (n / 48 == (int)rank)
                *reinterpret_cast<__nv_bfloat16 *>(B + swz32((uint32_t)((n % 48) * 32 + kk * 2))) = __float2bfloat16_rn(b_val(n, kk));
        } else {
            *reinterpret_cast<__nv_bfloat16 *>(B + swz32((uint32_t)(n * 32 + kk * 2))) = __float2bfloat16_rn(b_val(n, kk));
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc2(&slot, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t base = slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
    uint32_t ph = 0;
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            umma2_bf16(base, desc32(smem_u32(A)), desc32(smem_u32(B)), idesc, 0u);
            umma2_commit(&bar, 3);
        }
        __syncwarp();
    }
    mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    {
        const uint32_t t_addr = base + ((uint32_t)(32 * warp) << 16);
        uint32_t r0[32], r1[32], r2[32];
        tmem_ld32(t_addr, r0);
        tmem_ld32(t_addr + 32, r1);
        tmem_ld32(t_addr + 64, r2);
        tmem_ld_wait();
        float *o = out + ((size_t)rank * 128 + 32 * warp + lane) * N;
        for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(r0[j]); o[32 + j] = __uint_as_float(r1[j]); o[64 + j] = __uint_as_float(r2[j]); }
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    // ---- rate: one issuing thread of the leader, 8 UMMAs per iteration on 8 different A tiles, one accumulator ----
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const uint32_t a0 = smem_u32(A), b0 = smem_u32(B);
            const long long t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) umma2_bf16(base, desc32(a0 + (uint32_t)(u * 4096)), desc32(b0), idesc, 1u);
            }
            umma2_commit(&bar, 3);
            mbar_wait(&bar, ph);
            cycles[0] = clock64() - t0;
        }
        __syncwarp();
    }
    if (!(rank == 0 && warp == 0 )) mbar_wait(&bar, ph);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(base, 128);
}

int main()
{
    float *out;
    long long *cyc;
    cudaMalloc(&out, 2 * 128 * N * sizeof(float));
    cudaMalloc(&cyc, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 1920;
    for (int b_mode = 0; b_mode < 2; ++b_mode) {
        cudaMemset(out, 0, 2 * 128 * N * sizeof(float));
        cudaMemset(cyc, 0, 64);
        k<<<2, 128, 48 * 1024>>>(out, cyc, iters, b_mode);
        cudaError_t e = cudaDeviceSynchronize();
        static float h[2 * 128 * N];
        long long c = 0;
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // hypotheses for what column n of the accumulator of CTA c holds
        int ok_split = 0, ok_local = 0, total = 0;
        for (int cta = 0; cta < 2; ++cta)
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    float want = 0.f;
                    for (int kk = 0; kk < 16; ++kk) want += a_val(cta, r, kk) * b_val(n, kk);
                    const float got = h[((size_t)cta * 128 + r) * N + n];
                    ++total;
                    ok_split += got == want;
                    // "local" hypothesis: every CTA only sees its own smem rows 0..95 as B
                    float alt = 0.f;
                    if (b_mode == 0) {
                        const int src = n < 48 ? 48 * cta + n : -1;
                        if (src >= 0) for (int kk = 0; kk < 16; ++kk) alt += a_val(cta, r, kk) * b_val(src, kk);
                    }
                    ok_local += got == alt;
                }
        printf("b_mode %d (%s): %s | D == A x B^T with B = [rows of CTA 0 ; rows of CTA 1]: %d / %d | 'each CTA uses only its own rows': %d / %d | "
               "%.1f cycles per pair-UMMA (M = 256, N = %d, K = 16), one issuing thread\n",
               b_mode, b_mode == 0 ? "each CTA holds its 48 rows" : "both CTAs hold all 96 rows", cudaGetErrorString(e), ok_split, total,
               ok_local, total, (double)c / iters, N);
        printf("   sample D[cta0][r=5][n=0,1,47,48,95] = %g %g %g %g %g ; D[cta1][r=5][...] = %g %g %g %g %g\n", h[5 * N], h[5 * N + 1], h[5 * N + 47],
               h[5 * N + 48], h[5 * N + 95], h[(128 + 5) * N], h[(128 + 5) * N + 1], h[(128 + 5) * N + 47], h[(128 + 5) * N + 48], h[(128 + 5) * N + 95]);
    }
    return 0;
}
