// Probe: what limits ONE issuing thread to ~78 cycles per tcgen05.mma (umma_rate_probe.cu), and how fast the
// epilogue side can read accumulators back.
//   (a) rolled loop, 4 accumulators round-robin                   (the 78-cycle baseline)
//   (b) loop unrolled x8, descriptors precomputed, 4 accumulators (is it the uniform-register hand-over?)
//   (c) unrolled x8, ONE accumulator                              (dependent accumulate chain)
//   (d) unrolled x8, accumulator base moving by 32 columns        (overlapping N = 96 windows, the column-tile trunk)
//   (e) two issuing threads, unrolled, separate accumulators      (aggregate)
//   (f) tcgen05.ld 32x32b.x32 by 4 / 8 / 12 warps                 (aggregate cycles per warp-load of 4 KB)
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_issue_probe.bin umma_issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tacult_b200/csrc/tcgen05.cuh"

using namespace tc;

constexpr int SW = 64;

__device__ __forceinline__ uint64_t desc64(uint32_t saddr)
{
    constexpr uint32_t hi = ((8u * SW) >> 4) | (1u << 14) | (4u << 29);
    return ((uint64_t)hi << 32) | (uint64_t)(saddr >> 4);
}

template <int MODE>
__device__ __forceinline__ void issue_unrolled(uint32_t base, uint32_t a0, uint32_t b0, uint32_t idesc, int iters)
{
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t a = a0 + (uint32_t)(u * 128 * SW);
            uint32_t d;
            if (MODE == 0) d = base + (uint32_t)((u & 3) * 128);          // 4 accumulators
            else if (MODE == 1) d = base;                                  // one accumulator
            else d = base + (uint32_t)(32 * u);                            // overlapping windows
            umma_bf16(d, desc64(a), desc64(b0), idesc, 1u);
        }
    }
}

__global__ void k(long long *out, int n_cols, int iters, int ld_warps)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar, bars2[2];
    __shared__ long long t_end[16];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(bars2, 1); mbar_init(bars2 + 1, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
    const uint32_t idesc = make_instr_desc(n_cols);
    uint32_t ph = 0;
    if (threadIdx.x == 0) {
        for (int mode = -1; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                const long long t0 = clock64();
                if (mode < 0) {
                    for (int i = 0; i < iters; ++i) {
                        const uint32_t a = a0 + (uint32_t)((i & 7) * 128 * SW);
                        umma_bf16(base + (uint32_t)((i & 3) * 128), make_smem_desc<SW>(a), make_smem_desc<SW>(b0), idesc, 1u);
                    }
                } else if (mode == 0) issue_unrolled<0>(base, a0, b0, idesc, iters);
                else if (mode == 1) issue_unrolled<1>(base, a0, b0, idesc, iters);
                else issue_unrolled<2>(base, a0, b0, idesc, iters);
                umma_commit(&bar);
                mbar_wait(&bar, ph);
                ph ^= 1;
                out[mode + 1] = clock64() - t0;
            }
        }
    }
    __syncthreads();
    // (e) two issuers
    {
        const int w = threadIdx.x >> 5;
        const long long t0 = clock64();
        if ((threadIdx.x & 31) == 0 && w < 2) {
            issue_unrolled<0>(base + (uint32_t)(w * 32), a0 + (uint32_t)(w * 8 * 128 * SW), b0, idesc, iters);
            umma_commit(bars2 + w);
            mbar_wait(bars2 + w, 0);
            t_end[w] = clock64() - t0;
        }
        __syncthreads();
        if (threadIdx.x == 0) out[4] = max(t_end[0], t_end[1]);
    }
    // (f) accumulator read-back
    tc_fence_after();
    __syncthreads();
    {
        const int w = threadIdx.x >> 5;
        const long long t0 = clock64();
        uint32_t sink = 0;
        if (w < ld_warps) {
            const uint32_t t_addr = base + ((uint32_t)(32 * (w & 3)) << 16);
            for (int i = 0; i < iters; i += 3) {
                uint32_t r0[32], r1[32], r2[32];
                tmem_ld32(t_addr + (uint32_t)((i & 3) * 96), r0);
                tmem_ld32(t_addr + (uint32_t)((i & 3) * 96 + 32), r1);
                tmem_ld32(t_addr + (uint32_t)((i & 3) * 96 + 64), r2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sink ^= r0[j] + r1[j] + r2[j];
            }
            t_end[w] = clock64() - t0;
            if (sink == 0x12345678u) out[15] = sink;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long m = 0;
            for (int i = 0; i < ld_warps; ++i) m = max(m, t_end[i]);
            out[5] = m;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}


// (g) the UMMA stream of one layer of k_trunk_col, feature by feature: 9 tiles x (3 dy x 2 k-slices) UMMAs of N = 96.
//     flags: 1 commit after every tile, 2 A shifted by (dy-1) rows, 4 second k-slice at +32 B, 8 one B tile per (dy, k),
//     16 accumulator window moves by 32 columns per tile, 32 a bias UMMA (N = 32, overwrite) before each tile
template <int UNROLL>
__global__ void k2(long long *out, int flags, int reps)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar, tbar[16];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 16; ++i) mbar_init(tbar + i, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    const uint32_t x0 = smem_u32(smem) + 1024, w0 = smem_u32(smem + 100 * 1024), ones = smem_u32(smem + 130 * 1024);
    const uint32_t idesc96 = make_instr_desc(96), idesc32 = make_instr_desc(32);
    const int sa = (flags & 2) ? 64 : 0, sk = (flags & 4) ? 32 : 0, sb = (flags & 8) ? 1 : 0, sd = (flags & 16) ? 32 : 0;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        for (int rep = 0; rep < 2; ++rep) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
                for (int i = 0; i < 9; ++i) {
                    const uint32_t d = base + (uint32_t)(sd * i);
                    const uint32_t a_tile = x0 + (uint32_t)(i * 8192);
                    if (flags & 32) umma_bf16(d + 64, desc64(ones), desc64(w0 + 20 * 1024), idesc32, 0u);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_bf16(d, desc64(((flags & 64) ? x0 + (uint32_t)(((i * 6 + dy * 2 + k) & 7) * 8192) : a_tile) +
                                                (uint32_t)((dy - 1) * sa + k * sk)),
                                      desc64(w0 + (uint32_t)((dy * 6144 + k * 32) * sb)), idesc96, 1u);
                    if (flags & 1) umma_commit(tbar + i);
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, ph);
            ph ^= 1;
            out[rep] = clock64() - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 128);
    const int iters = 1920;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    for (int n : {32, 96}) {
        for (int ldw : {4, 8, 12}) {
            long long h[16] = {0};
            cudaMemset(d, 0, 128);
            k<<<1, 32 * 12, 165 * 1024>>>(d, n, iters, ldw);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("N = %3d: %s | one issuer, cycles per UMMA: rolled %.1f, unrolled x8 %.1f, one accumulator %.1f, overlapping "
                   "windows %.1f | two issuers (aggregate) %.1f | tcgen05.ld.x32 by %2d warps: %.1f cycles per warp-load "
                   "(aggregate), %.1f per warp\n",
                   n, cudaGetErrorString(e), (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters,
                   (double)h[3] / iters, (double)h[4] / (2.0 * iters), ldw, (double)h[5] / ((double)iters * ldw),
                   (double)h[5] / iters);
        }
    }
    cudaFuncSetAttribute(k2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaFuncSetAttribute(k2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaFuncSetAttribute(k2<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    for (int flags : {0, 64, 64 + 31, 64 + 63, 2048 + 64, 2048 + 64 + 31, 2048 + 64 + 63, 64 + 16, 64 + 8, 64 + 6}) {
        long long h[2] = {0, 0};
        const int reps = 40;
        if (flags & 2048) k2<9><<<1, 128, 165 * 1024>>>(d, flags, reps);
        else if (flags & 1024) k2<3><<<1, 128, 165 * 1024>>>(d, flags, reps);
        else k2<1><<<1, 128, 165 * 1024>>>(d, flags, reps);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("layer stream (+1024: tile loop unrolled x3, +2048: x9), flags %4d: %s  %.1f cycles per tile of 6 N=96 UMMAs%s\n", flags, cudaGetErrorString(e),
               (double)h[1] / (reps * 9.0), (flags & 32) ? " + 1 N=32" : "");
    }
    return 0;
}
