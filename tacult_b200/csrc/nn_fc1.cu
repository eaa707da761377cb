// fc1 (81*C -> E) with the reduction dimension split over a pair of CTAs (thread-block cluster of 2).
//
// Why: at self-play batch sizes fc1 is latency-bound, not bandwidth- or math-bound.  A [128 x 64] output tile streams
// 81*C/64 operand chunks of 24 KB through a ring that the shared-memory budget (a fused-trunk CTA of the other
// partition lives on the same SM) limits to 3 stages; with ~1.3 us per L2 round trip that is ~14 round trips per CTA
// (ncu: 23 us for 1024 rows, tensor pipe 15 % active, one CTA on 32..64 of the 148 SMs).  Splitting K over two CTAs
// doubles the bytes in flight and the SMs in use and halves the chain of round trips.  The partial accumulator of the
// second CTA goes to the first one through distributed shared memory (st.shared::cluster into the first CTA's idle
// operand ring), which adds bias, ReLU, packs to bf16 and stores the embedding rows.
//
// k_fc1_heads adds the policy / value heads (network.py:97-100) to the same launch.  The heads GEMM of a 128-row tile
// (K = E, N = 96) needs the embedding rows of all E/64 column tiles, which different clusters produce: every cluster
// publishes its slice to HBM/L2, bumps a per-row-tile arrival counter, and the cluster that arrives LAST runs the heads
// for that row tile on the spot — embedding tile and head weights stream through its drained operand ring, a second
// accumulator of 96 TMEM columns, and the soft-max (optionally masked with the leaf's legal moves and renormalised,
// mcts.py:272-275 in float32) + tanh epilogue.  That removes a kernel boundary (launch, drain, TMEM / barrier set-up:
// 12 us for a 12 MFLOP GEMM) from every simulation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "nn_bf16.cuh"
#include "rules.cuh"
#include "tcgen05.cuh"

namespace tc {

namespace fc1 {

constexpr int THREADS = 192;             // warp 0: TMA producer, warp 1: UMMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int BLOCK_N = 64, BLOCK_K = 64, SW = 128;
constexpr int A_BYTES = 128 * SW, B_BYTES = BLOCK_N * SW, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int MAX_STAGES = 8;

struct Params {
    const int *count_dev;
    int batch;
    int num_kb;              // ceil(81*C / 64)
    int n_blocks;            // E / 64
    int stages;
    int ldc;                 // E
    const float *bias;
    __nv_bfloat16 *out;      // [row][E]
};

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
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, uint32_t bits)
{
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(bits) : "memory");
}

__global__ void __launch_bounds__(THREADS) k_fc1_splitk(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB, Params P)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)P.stages * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *tfull = bars + 2 * MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();                                       // the trunk's output and the live count are read from here on

    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    const int n_tiles = ((count + 127) >> 7) * P.n_blocks;
    const int tile = blockIdx.x >> 1;                     // both CTAs of a cluster share the output tile
    const uint32_t rank = cluster_ctarank();
    if (tile < n_tiles) {                                 // uniform over the cluster
        const int mt = tile / P.n_blocks, nb = tile - mt * P.n_blocks;
        const int half = (P.num_kb + 1) >> 1;
        const int kb0 = rank ? half : 0, kb1 = rank ? P.num_kb : half;
        uint32_t a0[32], a1[32];                          // this thread's accumulator row (epilogue warps)
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    uint8_t *sa = smem + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(full + stage, (uint32_t)STAGE_BYTES);
                    tma_load_2d(sa, &tmA, full + stage, kb * BLOCK_K, 128 * mt);
                    tma_load_2d(sa + A_BYTES, &tmB, full + stage, kb * BLOCK_K, nb * BLOCK_N);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = make_instr_desc(BLOCK_N);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_bf16(tmem_base, make_smem_desc<SW>(a_addr + 32 * k), make_smem_desc<SW>(b_addr + 32 * k), idesc,
                                  (uint32_t)((kb != kb0) | k));
                    umma_commit(empty + stage);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull);
            }
        } else {
            mbar_wait(tfull, 0);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
            tmem_ld32(t_base, a0);
            tmem_ld32(t_base + 32u, a1);
            tmem_ld_wait();
            tc_fence_before();
        }
        __syncwarp();
        // (1) both accumulators are in registers and CTA 0's operand ring is drained (its epilogue warps have seen tfull)
        cluster_sync_all();
        const int quarter = warp & 3;
        const int r = 32 * quarter + lane;                // row inside the tile
        float *part = reinterpret_cast<float *>(smem);    // [64 columns][128 rows] in CTA 0
        if (warp >= 2 && rank == 1) {
            const uint32_t remote = map_to_cta(smem_u32(part), 0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                st_cluster_f32(remote + (uint32_t)((j * 128 + r) * 4), a0[j]);
                st_cluster_f32(remote + (uint32_t)(((32 + j) * 128 + r) * 4), a1[j]);
            }
        }
        // (2) the partial sums have landed in CTA 0
        cluster_sync_all();
        if (warp >= 2 && rank == 0) {
            const int q = 128 * mt + r;
            if (q < count) {
                const float *bias = P.bias + nb * BLOCK_N;
                __nv_bfloat16 *o = P.out + (size_t)q * P.ldc + nb * BLOCK_N;
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    uint4 ov;
                    __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = j + 2 * t;
                        const float x0 = __uint_as_float(c < 32 ? a0[c] : a1[c - 32]) + part[c * 128 + r] + __ldg(bias + c);
                        const float x1 = __uint_as_float(c + 1 < 32 ? a0[c + 1] : a1[c + 1 - 32]) + part[(c + 1) * 128 + r] +
                                         __ldg(bias + c + 1);
                        op[t] = __floats2bfloat162_rn(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
                    }
                    *reinterpret_cast<uint4 *>(o + j) = ov;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}


// ---- fc1 + heads in one launch --------------------------------------------------------------------------------------
constexpr int HB_BYTES = 96 * SW;                          // one K chunk of the head weights [96 x 64]
constexpr int STAGE_H = A_BYTES + HB_BYTES;                // ring stage of the fused kernel (fc1 uses 24 of its 28 KB)
constexpr int HEAD_COLS = 96;
constexpr int TMEM_COLS_H = 256;                           // fc1 accumulator [0, 64) + heads accumulator [64, 160)

struct HParams {
    const int *count_dev;
    int batch;
    int num_kb;              // ceil(81*C / 64)
    int n_blocks;            // E / 64
    int stages;
    int ldc;                 // E
    const float *bias;       // fc1 bias [E]
    __nv_bfloat16 *emb;      // [row][E] scratch the slices are exchanged through
    int *ready;              // [row tiles] arrival counters, zero between launches
    const float *head_bias;  // [96]
    float *pi;               // [row][81]
    float *v;                // [row]
    const uint32_t *states;  // packed leaf positions (masked mode), else null
    int masked;
    int pdl_trigger;
    long long *marks;        // optional (TACULT_FC1_TIMELINE): clock64 stamps of one epilogue thread of every cluster's CTA 0
};

__global__ void __launch_bounds__(THREADS) k_fc1_heads(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB,
                                                      const __grid_constant__ CUtensorMap tmE,
                                                      const __grid_constant__ CUtensorMap tmH, HParams P)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)P.stages * STAGE_H);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *tfull = bars + 2 * MAX_STAGES, *hfull = bars + 2 * MAX_STAGES + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 2);
    int *s_last = reinterpret_cast<int *>(tmem_slot + 1);

    griddep_launch(P.pdl_trigger == 1 || P.pdl_trigger == 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *mk = (P.marks && threadIdx.x == 64 && (blockIdx.x & 1) == 0 && (blockIdx.x >> 1) < 256)
                        ? P.marks + 16 * (blockIdx.x >> 1) : nullptr;
    if (mk) mk[0] = clock64();
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        mbar_init(hfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmE);
        tma_prefetch_desc(&tmH);
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS_H);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (mk) mk[1] = clock64();
    griddep_wait();
    if (mk) mk[2] = clock64();

    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    const int n_tiles = ((count + 127) >> 7) * P.n_blocks;
    const int tile = blockIdx.x >> 1;
    const uint32_t rank = cluster_ctarank();
    bool run_heads = false;
    int mt = 0, n_done = 0;
    if (tile < n_tiles) {                                 // uniform over the cluster
        mt = tile / P.n_blocks;
        const int nb = tile - mt * P.n_blocks;
        const int half = (P.num_kb + 1) >> 1;
        const int kb0 = rank ? half : 0, kb1 = rank ? P.num_kb : half;
        n_done = kb1 - kb0;
        uint32_t a0[32], a1[32];
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    uint8_t *sa = smem + (size_t)stage * STAGE_H;
                    mbar_expect_tx(full + stage, (uint32_t)STAGE_BYTES);
                    tma_load_2d(sa, &tmA, full + stage, kb * BLOCK_K, 128 * mt);
                    tma_load_2d(sa + A_BYTES, &tmB, full + stage, kb * BLOCK_K, nb * BLOCK_N);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = make_instr_desc(BLOCK_N);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_H);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_bf16(tmem_base, make_smem_desc<SW>(a_addr + 32 * k), make_smem_desc<SW>(b_addr + 32 * k), idesc,
                                  (uint32_t)((kb != kb0) | k));
                    umma_commit(empty + stage);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull);
            }
        } else {
            mbar_wait(tfull, 0);
            if (mk) mk[3] = clock64();
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
            tmem_ld32(t_base, a0);
            tmem_ld32(t_base + 32u, a1);
            tmem_ld_wait();
            tc_fence_before();
        }
        __syncwarp();
        cluster_sync_all();
        if (mk) mk[4] = clock64();
        const int quarter = warp & 3;
        const int r = 32 * quarter + lane;
        float *part = reinterpret_cast<float *>(smem);    // [64 columns][128 rows] in CTA 0 (32 KB of the drained ring)
        if (warp >= 2 && rank == 1) {
            const uint32_t remote = map_to_cta(smem_u32(part), 0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                st_cluster_f32(remote + (uint32_t)((j * 128 + r) * 4), a0[j]);
                st_cluster_f32(remote + (uint32_t)(((32 + j) * 128 + r) * 4), a1[j]);
            }
        }
        cluster_sync_all();
        if (mk) mk[5] = clock64();
        if (warp >= 2 && rank == 0) {
            const int q = 128 * mt + r;
            if (q < count) {
                const float *bias = P.bias + nb * BLOCK_N;
                __nv_bfloat16 *o = P.emb + (size_t)q * P.ldc + nb * BLOCK_N;
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    uint4 ov;
                    __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = j + 2 * t;
                        const float x0 = __uint_as_float(c < 32 ? a0[c] : a1[c - 32]) + part[c * 128 + r] + __ldg(bias + c);
                        const float x1 = __uint_as_float(c + 1 < 32 ? a0[c + 1] : a1[c + 1 - 32]) + part[(c + 1) * 128 + r] +
                                         __ldg(bias + c + 1);
                        op[t] = __floats2bfloat162_rn(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
                    }
                    *reinterpret_cast<uint4 *>(o + j) = ov;
                }
            }
        }
        if (rank == 0) {
            // publish this slice; whoever completes the row tile runs its heads
            if (mk) mk[6] = clock64();
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                const int old = atomicAdd(P.ready + mt, 1);
                const int last = old == P.n_blocks - 1;
                if (last) {
                    P.ready[mt] = 0;                      // ready for the next launch
                    __threadfence();
                }
                *s_last = last;
            }
            __syncthreads();
            run_heads = *s_last != 0;
            if (mk) { mk[7] = clock64(); mk[12] = run_heads; mk[13] = n_done; }
        }
    }
    // late trigger (TACULT_PDL_TRIGGER=2): the fc1 part of this CTA is over; the descent kernel's blocks may take the SMs
    // that CTAs without a heads tile free, and park in griddep_wait until the heads tiles are done
    if (threadIdx.x == 0) griddep_launch(P.pdl_trigger == 2);
    if (run_heads) {
        // ===== heads of row tile mt: [128 x E] (just written by E/64 clusters, L2 resident) x [96 x E]^T =====
        const int h_chunks = P.n_blocks;                  // E / 64
        int stage = n_done % P.stages;
        uint32_t phase = (uint32_t)(n_done / P.stages) & 1u;
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");      // the slices were written through the generic proxy
                for (int kc = 0; kc < h_chunks; ++kc) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    uint8_t *sa = smem + (size_t)stage * STAGE_H;
                    mbar_expect_tx(full + stage, (uint32_t)STAGE_H);
                    tma_load_2d(sa, &tmE, full + stage, kc * BLOCK_K, 128 * mt);
                    tma_load_2d(sa + A_BYTES, &tmH, full + stage, kc * BLOCK_K, 0);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = make_instr_desc(HEAD_COLS);
                for (int kc = 0; kc < h_chunks; ++kc) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_H);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_bf16(tmem_base + 64u, make_smem_desc<SW>(a_addr + 32 * k), make_smem_desc<SW>(b_addr + 32 * k),
                                  idesc, (uint32_t)(kc | k));
                    umma_commit(empty + stage);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(hfull);
            }
        } else {
            const int quarter = warp & 3;
            const int q = 128 * mt + 32 * quarter + lane;
            const bool valid = q < count;
            uint32_t am[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0x1FFFFu};
            if (P.masked && P.states && valid) {          // legal moves of this leaf, action order (mcts.py:272)
                uint32_t wd[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) wd[k] = __ldg(P.states + 8 * (size_t)q + k);
                Pos p = unpack(wd);
                uint32_t lw[3], m[3];
                legal_words(p, summarize(p), lw);
                to_action_order(lw, m);
                if (m[0] | m[1] | m[2]) { am[0] = m[0]; am[1] = m[1]; am[2] = m[2]; }
            }
            mbar_wait(hfull, 0);
            tc_fence_after();
            if (mk) mk[8] = clock64();
            const uint32_t t_base = tmem_base + ((uint32_t)(32 * quarter) << 16) + 64u;
            uint32_t h0[32], h1[32], h2[32];
            tmem_ld32(t_base, h0);
            tmem_ld32(t_base + 32u, h1);
            tmem_ld32(t_base + 64u, h2);
            tmem_ld_wait();
            // the operand ring is drained (hfull): its first 41 KB stage the rows for coalesced stores
            float *stage_rows = reinterpret_cast<float *>(smem);
            if (valid) {
                float lg[96];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    lg[j] = __uint_as_float(h0[j]) + __ldg(P.head_bias + j);
                    lg[32 + j] = __uint_as_float(h1[j]) + __ldg(P.head_bias + 32 + j);
                    lg[64 + j] = __uint_as_float(h2[j]) + __ldg(P.head_bias + 64 + j);
                }
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 81; ++j)
                    if ((am[j >> 5] >> (j & 31)) & 1u) mx = fmaxf(mx, lg[j]);
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 81; ++j) {
                    const bool on = (am[j >> 5] >> (j & 31)) & 1u;
                    lg[j] = on ? __expf(lg[j] - mx) : 0.f;
                    sum += lg[j];
                }
                const float inv = 1.f / sum;
                P.v[q] = tanhf(lg[81]);
                float *sr = stage_rows + (32 * quarter + lane) * 81;
#pragma unroll
                for (int j = 0; j < 81; ++j) sr[j] = lg[j] * inv;
            }
            __syncwarp();
            const int row0 = 128 * mt + 32 * quarter;
            const int n_rows = min(32, count - row0);
            if (n_rows > 0) {
                const float4 *src = reinterpret_cast<const float4 *>(stage_rows + 32 * quarter * 81);
                float4 *dstp = reinterpret_cast<float4 *>(P.pi + (size_t)row0 * 81);
                const int n4 = n_rows * 81 / 4;
                for (int i = lane; i < n4; i += 32) dstp[i] = src[i];
                const float *s1 = stage_rows + 32 * quarter * 81;
                float *d1 = P.pi + (size_t)row0 * 81;
                for (int i = 4 * n4 + lane; i < n_rows * 81; i += 32) d1[i] = s1[i];
            }
        }
    }
    if (mk) mk[9] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS_H);
}

}  // namespace fc1

size_t fc1_heads_smem(int stages) { return 1024 + (size_t)stages * fc1::STAGE_H + (2 * fc1::MAX_STAGES + 4) * 8 + 16; }

int launch_fc1_heads(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &emb_map, const CUtensorMap &head_w,
                     int batch, const int *count_dev, int K, int E, int stages, const float *bias, __nv_bfloat16 *emb,
                     int *ready, const float *head_bias, float *pi, float *v, const uint32_t *states, int masked,
                     cudaStream_t st)
{
    fc1::HParams P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.num_kb = (K + fc1::BLOCK_K - 1) / fc1::BLOCK_K;
    P.n_blocks = E / fc1::BLOCK_N;
    P.stages = stages;
    P.ldc = E;
    P.bias = bias;
    P.emb = emb;
    P.ready = ready;
    P.head_bias = head_bias;
    P.pi = pi;
    P.v = v;
    P.states = states;
    P.masked = masked;
    P.pdl_trigger = pdl_trigger_enabled();
    static long long *marks_dev = nullptr;
    if (getenv("TACULT_FC1_TIMELINE")) {
        if (!marks_dev) cudaMalloc(&marks_dev, 16 * 256 * sizeof(long long));
        cudaMemsetAsync(marks_dev, 0, 16 * 256 * sizeof(long long), st);
        P.marks = marks_dev;
    }
    const size_t smem = fc1_heads_smem(stages);
    static bool attr_set[64] = {false};
    int dev = 0;
    TC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        TC_CUDA(cudaFuncSetAttribute(fc1::k_fc1_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        prefer_max_shared(fc1::k_fc1_heads);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ((batch + 127) / 128) * P.n_blocks);
    cfg.blockDim = dim3(fc1::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    TC_CUDA(cudaLaunchKernelEx(&cfg, fc1::k_fc1_heads, a, b, emb_map, head_w, P));
    if (P.marks) {          // debugging aid: where a launch spends its time, per cluster (cycles since that CTA's entry)
        static long long h[16 * 256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, marks_dev, sizeof h, cudaMemcpyDeviceToHost);
        long long sum[10] = {0}, heads_n = 0, n = 0;
        for (int t = 0; t < 256; ++t) {
            const long long *m = h + 16 * t;
            if (m[0] == 0 || m[9] == 0) continue;
            ++n;
            for (int k = 1; k < 10; ++k) sum[k] += (m[k] ? m[k] - m[0] : 0);
            heads_n += m[12];
        }
        if (n)
            fprintf(stderr, "[fc1 marks] %lld clusters (%lld ran heads); mean cycles since CTA entry: prologue %lld | dep wait "
                    "over %lld | fc1 accumulators done %lld | cluster sync 1 %lld | partials landed %lld | slice stored "
                    "%lld | arrival known %lld | heads accumulator done (heads CTAs) %lld | exit (all CTAs) %lld\n", n,
                    heads_n, sum[1] / n, sum[2] / n, sum[3] / n, sum[4] / n, sum[5] / n, sum[6] / n, sum[7] / n,
                    heads_n ? sum[8] / heads_n : 0, sum[9] / n);
    }
    return TC_OK;
}

size_t fc1_splitk_smem(int stages) { return 1024 + (size_t)stages * fc1::STAGE_BYTES + (2 * fc1::MAX_STAGES + 2) * 8 + 16; }

int launch_fc1_splitk(const CUtensorMap &a, const CUtensorMap &b, int batch, const int *count_dev, int K, int E, int stages,
                      const float *bias, __nv_bfloat16 *out, cudaStream_t st)
{
    fc1::Params P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.num_kb = (K + fc1::BLOCK_K - 1) / fc1::BLOCK_K;
    P.n_blocks = E / fc1::BLOCK_N;
    P.stages = stages;
    P.ldc = E;
    P.bias = bias;
    P.out = out;
    const size_t smem = fc1_splitk_smem(stages);
    // the opt-in is per device: one flag per ordinal (several engines on different devices may share the process)
    static bool attr_set[64] = {false};
    int dev = 0;
    TC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        TC_CUDA(cudaFuncSetAttribute(fc1::k_fc1_splitk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ((batch + 127) / 128) * P.n_blocks);
    cfg.blockDim = dim3(fc1::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    TC_CUDA(cudaLaunchKernelEx(&cfg, fc1::k_fc1_splitk, a, b, P));
    return TC_OK;
}

}  // namespace tc
