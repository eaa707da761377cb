// Probe: what exactly does tcgen05.shift.cta_group::1.down move?  (lanes, column width, boundary rows)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shift_probe tmem_shift_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tacult_b200/csrc/tcgen05.cuh"

using namespace tc;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__global__ void probe(uint32_t *out, int lane_base, int col, int times)
{
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    const uint32_t mine = base + ((uint32_t)(32 * warp) << 16);
    uint32_t v[16];
    for (int c0 = 0; c0 < 64; c0 += 16) {
        for (int j = 0; j < 16; ++j) v[j] = threadIdx.x * 1000 + c0 + j;
        tmem_st16(mine + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        for (int k = 0; k < times; ++k)
            asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + ((uint32_t)lane_base << 16) + (uint32_t)col) : "memory");
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld16(mine + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 64 + c0 + j] = r[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 64);
}

// cycles for n back-to-back shifts of distinct 8-column strips, issue -> commit -> mbarrier completion
__global__ void shift_cost(long long *out, int n)
{
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        for (int rep = 0; rep < 4; ++rep) {
            const long long t0 = clock64();
            for (int k = 0; k < n; ++k)
                asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + (uint32_t)(8 * (k % 12))) : "memory");
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, ph);
            ph ^= 1;
            const long long t2 = clock64();
            out[2 * rep] = t1 - t0;
            out[2 * rep + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 128);
}

int main()
{
    {
        long long *dt, ht[8];
        cudaMalloc(&dt, sizeof ht);
        for (int n : {0, 1, 12, 48, 96}) {
            shift_cost<<<1, 128>>>(dt, n);
            cudaDeviceSynchronize();
            cudaMemcpy(ht, dt, sizeof ht, cudaMemcpyDeviceToHost);
            printf("shift_cost n=%2d: issue/total cycles:", n);
            for (int r = 0; r < 4; ++r) printf("  %lld/%lld", ht[2 * r], ht[2 * r + 1]);
            printf("\n");
        }
    }
    uint32_t *d, h[128 * 64];
    cudaMalloc(&d, sizeof h);
    const int cases[][3] = {{0, 8, 1}};
    for (auto &c : cases) {
        cudaMemset(d, 0, sizeof h);
        probe<<<1, 128>>>(d, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case lane_base=%d col=%d times=%d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        // report every (lane, col) whose value differs from the identity pattern, summarised per column
        for (int col = 0; col < 64; ++col) {
            int changed = 0, first = -1, last = -1;
            for (int l = 0; l < 128; ++l)
                if (h[l * 64 + col] != (uint32_t)(l * 1000 + col)) { ++changed; if (first < 0) first = l; last = l; }
            if (changed) {
                printf("  col %2d: %3d lanes changed (%d..%d);", col, changed, first, last);
                for (int l : {0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 96, 97, 126, 127})
                    printf(" L%d=%u", l, h[l * 64 + col]);
                printf("\n");
            }
        }
    }
    return 0;
}
