// bf16 tensor-core forward of the dual-head ResNet ("production mode": within 2e-2 of the reference net).
// Graph: /root/reference/src/tacult/network.py:81-102.
//
// Every 3x3 convolution of the trunk and fc1 run as tcgen05 GEMMs with fp32 accumulators in TMEM, operands
// staged in shared memory by TMA (cp.async.bulk.tensor, 32/64/128-byte swizzle), BN-fold + bias + residual +
// ReLU fused in the TMEM->register epilogue.
//
// Implicit GEMM without an im2col buffer: activations live in HBM as bf16 rows [row][C] in a padded, stacked
// layout — position b occupies 100 rows, cell (r,c) at row 16 + 100 b + 10 r + c; the 10th column of every
// board row and the 10 rows after each position are zero and never written.  A 3x3 tap (dy,dx) is then a pure
// row offset 10 dy + dx, so the A tile of a tap is ONE 2-D TMA box at a shifted row coordinate and the zero
// halo comes for free.  81 of every 100 GEMM rows are real outputs.
//
// Kernel anatomy (persistent, warp specialised, one elected thread per async role):
//   warp 0  TMA producer      smem ring:  full[s] / empty[s] mbarriers
//   warp 1  tcgen05.mma issue TMEM ring:  tmem_full[2] / tmem_empty[2]   (also allocates / frees TMEM)
//   warps 2-5 epilogue        tcgen05.ld 32 lanes x 32 columns per warp, one GEMM row per thread
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "tcgen05.cuh"

namespace tc {

constexpr int PAD0 = 16;            // zero rows in front of position 0 (>= 11, the largest tap offset)
constexpr int ROWS_PER_POS = 100;
constexpr int GEMM_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int SH_ROWS = 160;        // shifted variant: rows [g0-16, g0+144) of the activation per 128-row tile
constexpr int SH_LEAD = 16;

enum { OUT_PADDED_BF16 = 0, OUT_COMPACT_BF16 = 1, OUT_F32 = 2, OUT_PLAIN_BF16 = 3, OUT_HEADS = 4 };

struct GemmParams {
    const int *count_dev;     // live positions (may be null)
    int batch;                // launch bound on positions
    int rows_per_unit;        // GEMM rows per position: 100 (conv) or 1 (fc1)
    int taps;                 // 9 (conv) or 1
    int k_chunks;             // BLOCK_K chunks per tap
    int block_n;              // N per tile (multiple of 16, <= 256)
    int n_blocks;             // N tiles
    int a_row0;               // first A row of GEMM row 0 (PAD0 for conv, 0 for fc1)
    int b_tap_rows;           // rows of the weight tensor per tap
    int stages;
    int variant;              // 0: one TMA box per tap; 1/2: ONE box per tile + 9 shifted UMMA descriptors (conv, K<=64)
    int mode;
    int relu;
    int ldc;                  // output row length in elements (C or E)
    const float *bias;
    const __nv_bfloat16 *residual;   // padded layout, may be null
    void *out;
    float *out_v;             // OUT_HEADS: value head [row]
    float *out_logits;        // OUT_HEADS: optional raw policy logits [row][81]
};

__device__ __forceinline__ int live_count(const int *count_dev, int batch)
{
    int n = batch;
    if (count_dev) n = min(n, *count_dev);
    return n;
}

// ------------------------------------------------------------------------------------------------
// the GEMM / implicit-conv kernel
// ------------------------------------------------------------------------------------------------
template <int BLOCK_K, bool HEADS>
__global__ void __launch_bounds__(GEMM_THREADS) k_gemm_tc(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB, GemmParams P)
{
    constexpr int SW = BLOCK_K * 2;
    constexpr int A_BYTES = 128 * SW;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = (P.block_n * SW + 1023) & ~1023;
    const bool shifted = P.variant != 0;
    // classic: stages x [A 128 rows | B]; shifted: [9 weight taps, resident] then stages x [A 160 rows]
    const int stage_bytes = shifted ? SH_ROWS * SW : A_BYTES + b_bytes;
    const int w_bytes = shifted ? 9 * b_bytes : 0;
    uint8_t *stage0 = smem + w_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage0 + (size_t)P.stages * stage_bytes);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *tfull = bars + 2 * MAX_STAGES, *tempty = bars + 2 * MAX_STAGES + 2;
    uint64_t *wbar = bars + 2 * MAX_STAGES + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = P.taps * P.k_chunks;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * P.block_n) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlaps the tail of the previous kernel in the stream; its outputs (the live count, the A
    // operand) are only touched from here on
    griddep_wait();
    const int count = live_count(P.count_dev, P.batch);
    const int m_total = count * P.rows_per_unit;
    const int m_tiles = (m_total + 127) >> 7;
    const int n_tiles = m_tiles * P.n_blocks;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && shifted) {
            // weights of all 9 taps once per CTA, then one activation box (with halo rows) per tile
            if ((int)blockIdx.x < n_tiles) {
                mbar_expect_tx(wbar, (uint32_t)(9 * P.block_n * SW));
                for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem + tap * b_bytes, &tmB, wbar, 0, tap * P.b_tap_rows);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(empty + stage, phase ^ 1u);
                mbar_expect_tx(full + stage, (uint32_t)(SH_ROWS * SW));
                tma_load_2d(stage0 + (size_t)stage * stage_bytes, &tmA, full + stage, 0, P.a_row0 + 128 * tile - SH_LEAD);
                if (++stage == P.stages) { stage = 0; phase ^= 1u; }
            }
        } else if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int mt = tile / P.n_blocks, nb = tile - mt * P.n_blocks;
                const int a_row = P.a_row0 + 128 * mt;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / P.k_chunks, kc = kb - tap * P.k_chunks;
                    const int shift = P.taps == 9 ? (tap / 3 - 1) * 10 + (tap % 3 - 1) : 0;
                    mbar_wait(empty + stage, phase ^ 1u);
                    uint8_t *sa = stage0 + (size_t)stage * stage_bytes;
                    mbar_expect_tx(full + stage, (uint32_t)(A_BYTES + P.block_n * SW));
                    tma_load_2d(sa, &tmA, full + stage, kc * BLOCK_K, a_row + shift);
                    tma_load_2d(sa + A_BYTES, &tmB, full + stage, kc * BLOCK_K, tap * P.b_tap_rows + nb * P.block_n);
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && shifted) {
            const uint32_t idesc = make_instr_desc(P.block_n);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            if ((int)blockIdx.x < n_tiles) mbar_wait(wbar, 0);
            const uint32_t w_addr = smem_u32(smem);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(tempty + as, aphase ^ 1u);
                mbar_wait(full + stage, phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * P.block_n);
                const uint32_t a_base = smem_u32(stage0 + (size_t)stage * stage_bytes);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    // tap (dy,dx) = the same tile, 10*dy+dx rows further down: only the descriptor start moves
                    const int shift = SH_LEAD + (tap / 3 - 1) * 10 + (tap % 3 - 1);
                    const uint32_t a_addr = a_base + (uint32_t)(shift * SW);
                    const uint32_t b_addr = w_addr + (uint32_t)(tap * b_bytes);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                        uint64_t adesc = make_smem_desc<SW>(a_addr + 32 * k);
                        if (P.variant == 2) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;     // matrix base offset
                        umma_bf16(d_tmem, adesc, make_smem_desc<SW>(b_addr + 32 * k), idesc, (uint32_t)((tap | k) != 0));
                    }
                }
                umma_commit(empty + stage);
                umma_commit(tfull + as);
                if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        } else if (lane == 0) {
            const uint32_t idesc = make_instr_desc(P.block_n);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(tempty + as, aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * P.block_n);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage0 + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_bf16(d_tmem, make_smem_desc<SW>(a_addr + 32 * k), make_smem_desc<SW>(b_addr + 32 * k), idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(empty + stage);           // smem slot is free once these MMAs have read it
                    if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull + as);                  // accumulator complete
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias / residual / ReLU -> HBM =====
        const int quarter = warp & 3;                     // TMEM lanes [32*quarter, +32) belong to this warp
        int as = 0;
        uint32_t aphase = 0;
        // HEADS: when every CTA has at most one tile the operand ring is idle once the accumulator is complete, and
        // its first 41 KB stage the soft-max rows for coalesced stores (alignment: row0 * 81 * 4 is a multiple of 16
        // because row0 is a multiple of 32)
        float *stage_rows = (HEADS && (int)gridDim.x >= n_tiles && !P.out_logits && 128 * 81 * 4 <= P.stages * stage_bytes)
                                ? reinterpret_cast<float *>(stage0) : nullptr;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mt = tile / P.n_blocks, nb = tile - mt * P.n_blocks;
            const int n0 = nb * P.block_n;
            const int q = 128 * mt + 32 * quarter + lane;      // GEMM row of this thread
            bool valid;
            size_t out_row;
            if (P.mode == OUT_F32 || P.mode == OUT_PLAIN_BF16 || P.mode == OUT_HEADS) {
                valid = q < m_total;
                out_row = (size_t)q;
            } else {
                const int b = q / ROWS_PER_POS, rem = q - b * ROWS_PER_POS;
                const int r = rem / 10, c = rem - 10 * r;
                valid = b < count && rem < 90 && c != 9;
                out_row = P.mode == OUT_PADDED_BF16 ? (size_t)(PAD0 + q) : (size_t)(b * 81 + r * 9 + c);
            }
            mbar_wait(tfull + as, aphase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(as * P.block_n);
            if constexpr (HEADS) {
                // policy + value heads: N = 96 columns (81 logits, the value pre-activation, padding).  The whole row
                // sits in this thread's registers, so the unmasked softmax (network.py:99) needs no shuffles.
                uint32_t a0[32], a1[32], a2[32];
                tmem_ld32(t_base, a0);
                tmem_ld32(t_base + 32u, a1);
                tmem_ld32(t_base + 64u, a2);
                tmem_ld_wait();
                if (valid) {
                    float lg[96];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        lg[j] = __uint_as_float(a0[j]) + __ldg(P.bias + j);
                        lg[32 + j] = __uint_as_float(a1[j]) + __ldg(P.bias + 32 + j);
                        lg[64 + j] = __uint_as_float(a2[j]) + __ldg(P.bias + 64 + j);
                    }
                    float mx = lg[0];
#pragma unroll
                    for (int j = 1; j < 81; ++j) mx = fmaxf(mx, lg[j]);
                    if (P.out_logits) {
                        float *lo = P.out_logits + out_row * 81;
#pragma unroll
                        for (int j = 0; j < 81; ++j) lo[j] = lg[j];
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 81; ++j) {
                        lg[j] = __expf(lg[j] - mx);
                        sum += lg[j];
                    }
                    const float inv = 1.f / sum;
                    P.out_v[out_row] = tanhf(lg[81]);
                    if (stage_rows) {
                        // row-strided 4-byte stores cost one LSU transaction per lane (measured: the store queue was
                        // the epilogue's main stall); go through shared memory and write the warp's 32 x 81 block,
                        // which is contiguous in the output, with 16-byte coalesced stores.  Pitch 81 words is odd,
                        // so the column-wise writes are conflict-free.
                        float *sr = stage_rows + (32 * quarter + lane) * 81;
#pragma unroll
                        for (int j = 0; j < 81; ++j) sr[j] = lg[j] * inv;
                    } else {
                        float *po = reinterpret_cast<float *>(P.out) + out_row * 81;
#pragma unroll
                        for (int j = 0; j < 81; ++j) po[j] = lg[j] * inv;
                    }
                }
                if (stage_rows) {
                    __syncwarp();
                    const int row0 = 128 * mt + 32 * quarter;
                    const int n_rows = min(32, m_total - row0);
                    if (n_rows > 0) {
                        const float4 *src = reinterpret_cast<const float4 *>(stage_rows + 32 * quarter * 81);
                        float4 *dstp = reinterpret_cast<float4 *>(reinterpret_cast<float *>(P.out) + (size_t)row0 * 81);
                        const int n4 = n_rows * 81 / 4;          // 32 rows x 81 words = 648 float4; ragged tails below
                        for (int i = lane; i < n4; i += 32) dstp[i] = src[i];
                        const float *s1 = stage_rows + 32 * quarter * 81;
                        float *d1 = reinterpret_cast<float *>(P.out) + (size_t)row0 * 81;
                        for (int i = 4 * n4 + lane; i < n_rows * 81; i += 32) d1[i] = s1[i];
                    }
                    __syncwarp();
                }
            } else {
            const int cw = P.block_n < 32 ? 16 : 32;
            for (int c0 = 0; c0 < P.block_n; c0 += cw) {
                uint32_t acc[32];
                if (cw == 32) tmem_ld32(t_base + (uint32_t)c0, acc); else tmem_ld16(t_base + (uint32_t)c0, acc);
                tmem_ld_wait();
                if (valid) {
                    const float *bias = P.bias + n0 + c0;
                    if (P.mode == OUT_F32) {
                        float *o = reinterpret_cast<float *>(P.out) + out_row * P.ldc + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (j >= cw) break;
                            float4 v;
                            v.x = __uint_as_float(acc[j]) + __ldg(bias + j);
                            v.y = __uint_as_float(acc[j + 1]) + __ldg(bias + j + 1);
                            v.z = __uint_as_float(acc[j + 2]) + __ldg(bias + j + 2);
                            v.w = __uint_as_float(acc[j + 3]) + __ldg(bias + j + 3);
                            if (P.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            *reinterpret_cast<float4 *>(o + j) = v;
                        }
                    } else {
                        __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(P.out) + out_row * P.ldc + n0 + c0;
                        const __nv_bfloat16 *res = (P.residual && P.mode != OUT_PLAIN_BF16)
                                                       ? P.residual + (size_t)(PAD0 + q) * P.ldc + n0 + c0 : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (j >= cw) break;
                            float v[8];
#pragma unroll
                            for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(acc[j + t]) + __ldg(bias + j + t);
                            if (res) {
                                uint4 rv = *reinterpret_cast<const uint4 *>(res + j);
                                const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    float2 f = __bfloat1622float2(rp[t]);
                                    v[2 * t] += f.x;
                                    v[2 * t + 1] += f.y;
                                }
                            }
                            uint4 ov;
                            __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                float a = v[2 * t], b2 = v[2 * t + 1];
                                if (P.relu) { a = fmaxf(a, 0.f); b2 = fmaxf(b2, 0.f); }
                                op[t] = __floats2bfloat162_rn(a, b2);
                            }
                            *reinterpret_cast<uint4 *>(o + j) = ov;
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + as);
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// conv_in (in_planes -> C) + folded BN + ReLU on CUDA cores (K = 36: not worth a tensor tile), writing the
// padded bf16 layout (or the compact one when the net has no residual block).  One block per position.
__global__ void __launch_bounds__(256) k_stem_bf16(int batch, const int *__restrict__ count_dev,
                                                   const uint32_t *__restrict__ states, const float *__restrict__ obs,
                                                   int in_planes, int C, const float *__restrict__ w,
                                                   const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out,
                                                   int compact)
{
    extern __shared__ float s_in[];        // [in_planes][121], zero border
    const int row = blockIdx.x;
    if (row >= live_count(count_dev, batch)) return;
    for (int i = threadIdx.x; i < in_planes * 121; i += blockDim.x) s_in[i] = 0.f;
    __syncthreads();
    if (states) {
        if (threadIdx.x < 81) {
            uint32_t wd[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) wd[k] = states[8 * (size_t)row + k];
            Pos p = unpack(wd);
            Summary s = summarize(p);
            uint32_t lw[3];
            legal_words(p, s, lw);
            int a = threadIdx.x, cell = (a / 9 + 1) * 11 + a % 9 + 1;
            bool mine = action_bit(p.mine, a), occ = action_bit(p.occ, a);
            s_in[cell] = mine ? 1.f : 0.f;
            s_in[121 + cell] = (occ && !mine) ? 1.f : 0.f;
            s_in[242 + cell] = action_bit(lw, a) ? 1.f : 0.f;
            s_in[363 + cell] = 1.f;
        }
    } else {
        for (int i = threadIdx.x; i < in_planes * 81; i += blockDim.x) {
            int pl = i / 81, a = i % 81;
            s_in[pl * 121 + (a / 9 + 1) * 11 + a % 9 + 1] = obs[(size_t)row * in_planes * 81 + i];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 81 * C; idx += blockDim.x) {
        int p = idx / C, c = idx % C;
        int base = (p / 9) * 11 + p % 9;
        float acc = bias[c];
        for (int pl = 0; pl < in_planes; ++pl) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
                acc = fmaf(s_in[pl * 121 + base + (t / 3) * 11 + t % 3], w[(t * in_planes + pl) * C + c], acc);
        }
        size_t orow = compact ? (size_t)row * 81 + p : (size_t)PAD0 + (size_t)row * ROWS_PER_POS + (p / 9) * 10 + p % 9;
        out[orow * C + c] = __float2bfloat16_rn(fmaxf(acc, 0.f));
    }
}

// Observation planes as a PC-channel (16 or 32) bf16 activation in the padded layout (channels >= in_planes are zero),
// so that conv_in runs through the same tensor-core implicit GEMM as the trunk (K = PC per tap).  One warp per position.
// 32 channels hold the 8-frame history variant (8 x 4 planes, BASELINE.json configs[4]); the reference has 4 planes
// (network.py:55).
template <int PC>
__global__ void __launch_bounds__(256) k_encode_planes_bf16(int batch, const int *__restrict__ count_dev,
                                                            const uint32_t *__restrict__ states,
                                                            const float *__restrict__ obs, int in_planes,
                                                            __nv_bfloat16 *__restrict__ out)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= live_count(count_dev, batch)) return;
    Pos p;
    uint32_t lw[3] = {0, 0, 0};
    if (states) {
        uint32_t wd[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) wd[k] = states[8 * (size_t)row + k];
        p = unpack(wd);
        legal_words(p, summarize(p), lw);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int a = lane + 32 * q;
        if (a >= 81) continue;
        float v[PC];
#pragma unroll
        for (int c = 0; c < PC; ++c) v[c] = 0.f;
        if (states) {
            const bool mine = action_bit(p.mine, a), occ = action_bit(p.occ, a);
            v[0] = mine ? 1.f : 0.f;
            v[1] = (occ && !mine) ? 1.f : 0.f;
            v[2] = action_bit(lw, a) ? 1.f : 0.f;
            v[3] = 1.f;
        } else {
#pragma unroll
            for (int c = 0; c < PC; ++c)
                if (c < in_planes) v[c] = obs[((size_t)row * in_planes + c) * 81 + a];
        }
        uint4 o[PC / 8];
        __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(o);
#pragma unroll
        for (int c = 0; c < PC / 2; ++c) op[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
        uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)PAD0 + (size_t)row * ROWS_PER_POS + (a / 9) * 10 + a % 9) * PC);
#pragma unroll
        for (int c = 0; c < PC / 8; ++c) dst[c] = o[c];
    }
}

// debug / test helper: padded or compact bf16 activations -> f32 [row][81][C]
__global__ void k_unpad_to_f32(int batch, int C, const __nv_bfloat16 *__restrict__ in, int compact, float *__restrict__ out)
{
    size_t n = (size_t)batch * 81 * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        size_t bp = i / C;
        int p = (int)(bp % 81);
        size_t b = bp / 81;
        size_t row = compact ? bp : (size_t)PAD0 + b * ROWS_PER_POS + (p / 9) * 10 + p % 9;
        out[i] = __bfloat162float(in[row * C + c]);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor [rows][cols] (cols contiguous), box [box_rows][box_cols], swizzle = box_cols * 2 bytes
static int make_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols)
{
    EncodeTiledFn fn = encode_tiled_fn();
    TC_CHECK(fn != nullptr, TC_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TC_CHECK(r == CUDA_SUCCESS, TC_ECUDA, "cuTensorMapEncodeTiled failed with code %d (rows %llu cols %llu box %u x %u)",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return TC_OK;
}

struct NetBF16Impl {
    tc_net_desc d{};
    int capacity = 0;
    int block_k = 0;            // conv K chunk (channels): 16, 32 or 64
    size_t act_rows = 0;
    DevBuf<__nv_bfloat16> act[3];          // padded layout
    DevBuf<__nv_bfloat16> flat;            // compact [cap][81*C], fc1 input
    DevBuf<float> emb;                     // [cap][E]   (fallback heads path)
    DevBuf<__nv_bfloat16> emb16;           // [cap][E]   fc1 output feeding the tensor-core heads
    DevBuf<__nv_bfloat16> head_w16;        // [96][E]: 81 policy rows, the value row, zero padding
    DevBuf<float> head_b96;
    CUtensorMap map_emb16, map_head_w;
    bool heads_on_tensor = false;
    DevBuf<float> stem_w, stem_b;
    DevBuf<__nv_bfloat16> planes;          // padded layout, stem_k channels (observation planes, zero padded)
    int stem_k = 16;                       // 16 (up to 16 planes) or 32 (the 8-frame history variant, 32 planes)
    DevBuf<__nv_bfloat16> stem_w16;        // [9][C][16]
    CUtensorMap map_planes, map_stem_w;
    bool stem_on_tensor = false;
    std::vector<DevBuf<__nv_bfloat16>> conv_w;   // [9][Cout][Cin]
    std::vector<DevBuf<float>> conv_b;
    DevBuf<__nv_bfloat16> fc1_w;           // [E][81*C], k = p*C + c
    DevBuf<float> fc1_b, head_w, head_b;
    CUtensorMap map_act[3], map_flat, map_fc1;
    CUtensorMap map_act_sh[3], map_planes_sh;      // 160-row boxes for the shifted-descriptor variant
    int conv_variant = 0;
    // fused shared-memory-resident trunk (C <= 64)
    bool fused = false;
    int fused_G = 0, fused_T = 0;
    // column-tiled fused trunk (C == 32): weights as [layer][ky][3C x K] blocks
    bool col = false;
    DevBuf<__nv_bfloat16> col_w_stem, col_w_conv;
    DevBuf<float> col_bias;                // [1 + L][C]
    CUtensorMap map_col_stem, map_col_conv;
    DevBuf<__nv_bfloat16> conv_w_all;      // [n_conv][9][C][C]
    DevBuf<__nv_bfloat16> bias_tiles;      // [1 + n_conv][C][16]: bias as a K-major UMMA operand (hi, lo, 0...)
    CUtensorMap map_conv_w_all, map_bias_tiles;
    std::vector<CUtensorMap> map_conv_w;
    int conv_stages = 4, fc1_stages = 4, fc1_block_n = 64;
    bool fc1_split_k = true;               // K split over a CTA pair (nn_fc1.cu); TACULT_FC1_SPLITK=0 disables
    bool fc1_heads = false;                // fc1 and the heads in one launch (k_fc1_heads); TACULT_FC1_HEADS=0 disables
    int fc1_heads_stages = 3;
    DevBuf<int> tile_ready;                // arrival counter per 128-row tile, zero between launches
    size_t side_smem_budget = 96 * 1024;       // what an fc1 / heads CTA may use next to a resident fused-trunk CTA
    int sm_count = SM_COUNT;
};

NetBF16::~NetBF16() { delete impl; }

static size_t gemm_smem_bytes(int block_k, int block_n, int stages, int variant = 0)
{
    size_t a = 128 * block_k * 2, b = ((size_t)block_n * block_k * 2 + 1023) & ~(size_t)1023;
    if (variant != 0) return 1024 + 9 * b + (size_t)stages * SH_ROWS * block_k * 2 + (2 * MAX_STAGES + 6) * 8 + 16;
    return 1024 + stages * (a + b) + (2 * MAX_STAGES + 6) * 8 + 16;
}

static int pick_stages(int block_k, int block_n, size_t budget, int variant = 0)
{
    int s = MAX_STAGES;
    while (s > 2 && gemm_smem_bytes(block_k, block_n, s, variant) > budget) --s;
    return s;
}

template <typename T>
static int upload(DevBuf<T> &dst, const std::vector<T> &src)
{
    TC_CUDA(dst.alloc(src.size()));
    TC_CUDA(cudaMemcpy(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TC_OK;
}

static int upload_impl(NetBF16Impl &m, const FoldedNet &f, int cap, bool share_sm)
{
    m.d = f.d;
    m.capacity = cap;
    const int C = f.d.channels, E = f.d.embedding_size, I = f.d.in_planes;
    m.block_k = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
    const char *var = getenv("TACULT_CONV_VARIANT");
    m.conv_variant = var ? atoi(var) : 1;      // measured on B200: variant 1 is exact, variant 2 (base offset) is wrong
    if (C > 64) m.conv_variant = 0;            // weights of all taps must fit in shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m.sm_count, cudaDevAttrMultiProcessorCount, dev);

    std::vector<float> t((size_t)9 * I * C);
    for (int co = 0; co < C; ++co)
        for (int ci = 0; ci < I; ++ci)
            for (int tap = 0; tap < 9; ++tap) t[((size_t)tap * I + ci) * C + co] = f.stem_w[((size_t)co * I + ci) * 9 + tap];
    TC_TRY(upload(m.stem_w, t));
    TC_TRY(upload(m.stem_b, f.stem_b));

    const size_t L = f.conv_w.size();
    m.conv_w.clear();
    m.conv_b.clear();
    m.conv_w.resize(L);
    m.conv_b.resize(L);
    m.map_conv_w.resize(L);
    std::vector<__nv_bfloat16> wb((size_t)9 * C * C);
    for (size_t l = 0; l < L; ++l) {
        for (int tap = 0; tap < 9; ++tap)
            for (int co = 0; co < C; ++co)
                for (int ci = 0; ci < C; ++ci)
                    wb[((size_t)tap * C + co) * C + ci] = __float2bfloat16_rn(f.conv_w[l][((size_t)co * C + ci) * 9 + tap]);
        TC_TRY(upload(m.conv_w[l], wb));
        TC_TRY(upload(m.conv_b[l], f.conv_b[l]));
        TC_TRY(make_map(&m.map_conv_w[l], m.conv_w[l].p, (uint64_t)9 * C, C, (uint32_t)C, (uint32_t)m.block_k));
    }
    // the column-tiled trunk (nn_trunk_col.cu) is the only producer of the fc1 input when it is on, and writes it with
    // the cells of a position in column-major order
    {
        const char *fz = getenv("TACULT_FUSED"), *cz = getenv("TACULT_TRUNK_COL");
        // A pass of the column-tiled trunk costs the same for 1 or 12 positions per SM (9 tiles x 7 layers x ~430 cycles); the
        // tap-per-UMMA trunk costs ~840 cycles per layer and 128 rows (1.28 positions).  Below ~6 positions per SM - engines
        // of up to 768 rows, e.g. the arena's 512 boards - the older kernel is the faster one (leaf-evaluation sweep at batch
        // 256: 7.3 vs 6.3 M evals/s; at 1024: 21.2 vs 24.0).  The choice is per engine (its capacity), never per call: a
        // row's result does not depend on the batch it is evaluated in.  TACULT_TRUNK_COL=2 forces the column trunk.
        const bool big_enough = cap > 768 || (cz && atoi(cz) == 2);
        m.col = trunk_col_supported(C, I) && !share_sm && (!fz || atoi(fz) != 0) && (!cz || atoi(cz) != 0) && big_enough;
    }
    // fc1: [E][c*81+p] (NCHW flatten, network.py:92) -> [E][p'*C+c] to match the NHWC activations, p' = p or its transpose
    std::vector<__nv_bfloat16> fb((size_t)E * 81 * C);
    for (int e = 0; e < E; ++e)
        for (int c = 0; c < C; ++c)
            for (int p = 0; p < 81; ++p) {
                const int pp = m.col ? (p % 9) * 9 + p / 9 : p;
                fb[(size_t)e * 81 * C + (size_t)pp * C + c] = __float2bfloat16_rn(f.fc1_w[(size_t)e * 81 * C + (size_t)c * 81 + p]);
            }
    TC_TRY(upload(m.fc1_w, fb));
    TC_TRY(upload(m.fc1_b, f.fc1_b));
    std::vector<float> hw((size_t)E * 82), hb(f.pi_b);
    for (int k = 0; k < E; ++k) {
        for (int o = 0; o < 81; ++o) hw[(size_t)k * 82 + o] = f.pi_w[(size_t)o * E + k];
        hw[(size_t)k * 82 + 81] = f.v_w[k];
    }
    hb.push_back(f.v_b[0]);
    TC_TRY(upload(m.head_w, hw));
    TC_TRY(upload(m.head_b, hb));

    m.act_rows = (size_t)PAD0 + (size_t)cap * ROWS_PER_POS + 128 + 32;
    for (int i = 0; i < 3; ++i) {
        TC_CUDA(m.act[i].alloc(m.act_rows * C));
        TC_CUDA(cudaMemset(m.act[i].p, 0, m.act[i].bytes()));       // pad rows stay zero for ever
        TC_TRY(make_map(&m.map_act[i], m.act[i].p, m.act_rows, C, 128, (uint32_t)m.block_k));
        if (C <= 64) TC_TRY(make_map(&m.map_act_sh[i], m.act[i].p, m.act_rows, C, SH_ROWS, (uint32_t)m.block_k));
    }
    m.stem_on_tensor = I <= 32;
    m.stem_k = I <= 16 ? 16 : 32;
    if (m.stem_on_tensor) {
        const int SK = m.stem_k;
        std::vector<__nv_bfloat16> sw((size_t)9 * C * SK, __float2bfloat16_rn(0.f));
        for (int tap = 0; tap < 9; ++tap)
            for (int co = 0; co < C; ++co)
                for (int ci = 0; ci < I; ++ci)
                    sw[((size_t)tap * C + co) * SK + ci] = __float2bfloat16_rn(f.stem_w[((size_t)co * I + ci) * 9 + tap]);
        TC_TRY(upload(m.stem_w16, sw));
        TC_CUDA(m.planes.alloc(m.act_rows * SK));
        TC_CUDA(cudaMemset(m.planes.p, 0, m.planes.bytes()));
        TC_TRY(make_map(&m.map_planes, m.planes.p, m.act_rows, SK, 128, SK));
        if (C <= 64) TC_TRY(make_map(&m.map_planes_sh, m.planes.p, m.act_rows, SK, SH_ROWS, SK));
        TC_TRY(make_map(&m.map_stem_w, m.stem_w16.p, (uint64_t)9 * C, SK, (uint32_t)C, SK));
    }
    const size_t flat_rows = std::max(cap, 128);
    TC_CUDA(m.flat.alloc(flat_rows * 81 * C));
    TC_CUDA(cudaMemset(m.flat.p, 0, m.flat.bytes()));
    TC_CUDA(m.emb.alloc((size_t)cap * E));
    m.heads_on_tensor = E % 64 == 0;
    if (m.heads_on_tensor) {
        const size_t emb_rows = std::max(cap, 128);
        TC_CUDA(m.emb16.alloc(emb_rows * E));
        TC_CUDA(cudaMemset(m.emb16.p, 0, m.emb16.bytes()));
        std::vector<__nv_bfloat16> hw16((size_t)96 * E, __float2bfloat16_rn(0.f));
        std::vector<float> hb96(96, 0.f);
        for (int o = 0; o < 81; ++o) {
            for (int k = 0; k < E; ++k) hw16[(size_t)o * E + k] = __float2bfloat16_rn(f.pi_w[(size_t)o * E + k]);
            hb96[o] = f.pi_b[o];
        }
        for (int k = 0; k < E; ++k) hw16[(size_t)81 * E + k] = __float2bfloat16_rn(f.v_w[k]);
        hb96[81] = f.v_b[0];
        TC_TRY(upload(m.head_w16, hw16));
        TC_TRY(upload(m.head_b96, hb96));
        TC_TRY(make_map(&m.map_emb16, m.emb16.p, (uint64_t)emb_rows, (uint64_t)E, 128, 64));
        TC_TRY(make_map(&m.map_head_w, m.head_w16.p, 96, (uint64_t)E, 96, 64));
    }
    {
        const char *fb = getenv("TACULT_FC1_BLOCK_N");
        m.fc1_block_n = E % 64 == 0 ? 64 : (E % 32 == 0 ? 32 : 16);
        if (fb && atoi(fb) > 0 && E % atoi(fb) == 0 && atoi(fb) % 32 == 0 && atoi(fb) <= 256) m.fc1_block_n = atoi(fb);
        const char *sk = getenv("TACULT_FC1_SPLITK");
        m.fc1_split_k = !sk || atoi(sk) != 0;
    }
    {
        const char *fh = getenv("TACULT_FC1_HEADS");
        // single-partition engines only: the fused kernel takes half of an SM's TMEM columns and ~87 KB of shared memory
        m.fc1_heads = (!fh || atoi(fh) != 0) && !share_sm && m.fc1_split_k && m.heads_on_tensor && m.fc1_block_n == 64;
        if (m.fc1_heads) {
            TC_CUDA(m.tile_ready.alloc((size_t)(cap + 127) / 128 + 1));
            TC_CUDA(cudaMemset(m.tile_ready.p, 0, m.tile_ready.bytes()));
            m.fc1_heads_stages = 3;
            if (const char *fs = getenv("TACULT_FC1_HEADS_STAGES"))
                if (atoi(fs) >= 2 && atoi(fs) <= 6) m.fc1_heads_stages = atoi(fs);
        }
    }
    TC_TRY(make_map(&m.map_flat, m.flat.p, (uint64_t)flat_rows, (uint64_t)81 * C, 128, 64));
    TC_TRY(make_map(&m.map_fc1, m.fc1_w.p, (uint64_t)E, (uint64_t)81 * C, (uint32_t)m.fc1_block_n, 64));

    {
        const char *fz = getenv("TACULT_FUSED");
        int G = 0, T = 0;
        m.fused = m.stem_on_tensor && I <= 16 && (!fz || atoi(fz) != 0) && fused_plan(C, G, T);
        TC_CHECK(m.fused || !m.col, TC_ESTATE, "column-tiled trunk selected but the fused path is unavailable (fc1 weights are already in its order)");
        if (m.fused) {
            // G, T are the largest group that fits.  Size the kernel for THIS instance's capacity instead: the group the
            // kernel would pick for a full batch (it re-balances from the live leaf count, never above these maxima), so
            // that the shared memory left on each SM lets the fc1 / heads CTAs of the other partition co-reside with a
            // trunk CTA instead of queueing behind the whole trunk grid.
            const char *gz = getenv("TACULT_FUSED_GMAX");
            if (gz && atoi(gz) > 0) G = std::min(G, atoi(gz));
            const int waves = std::max(1, (cap + m.sm_count * G - 1) / (m.sm_count * G));
            G = std::max(1, std::min(G, (cap + m.sm_count * waves - 1) / (m.sm_count * waves)));
            T = (100 * G + 127) / 128;
            m.fused_G = G;
            m.fused_T = T;
            // only an instance that shares the SMs with other partitions' kernels limits its fc1 / heads CTAs to the
            // shared memory a trunk CTA leaves free
            if (share_sm) m.side_smem_budget = (size_t)227 * 1024 - fused_smem_bytes(C, T) - 2048;
            std::vector<__nv_bfloat16> wall(std::max<size_t>(1, L) * 9 * C * C);
            std::vector<float> ball((1 + L) * (size_t)C);
            for (int c = 0; c < C; ++c) ball[c] = f.stem_b[c];
            for (size_t l = 0; l < L; ++l) {
                for (int tap = 0; tap < 9; ++tap)
                    for (int co = 0; co < C; ++co)
                        for (int ci = 0; ci < C; ++ci)
                            wall[((l * 9 + tap) * C + co) * C + ci] =
                                __float2bfloat16_rn(f.conv_w[l][((size_t)co * C + ci) * 9 + tap]);
                for (int c = 0; c < C; ++c) ball[(1 + l) * C + c] = f.conv_b[l][c];
            }
            std::vector<__nv_bfloat16> btiles;
            fused_bias_tiles(ball, btiles);
            TC_TRY(upload(m.conv_w_all, wall));
            TC_TRY(upload(m.bias_tiles, btiles));
            TC_TRY(make_map(&m.map_bias_tiles, m.bias_tiles.p, (uint64_t)(1 + L) * C, 16, (uint32_t)C, 16));
            TC_TRY(make_map(&m.map_conv_w_all, m.conv_w_all.p, std::max<size_t>(1, L) * 9 * C, C, (uint32_t)C, (uint32_t)C));
            if (m.col) {
                std::vector<__nv_bfloat16> ws((size_t)9 * C * 16), wc(std::max<size_t>(1, L) * 9 * C * C, __float2bfloat16_rn(0.f));
                trunk_col_pack(f.stem_w.data(), C, I, 16, ws.data());
                for (size_t l = 0; l < L; ++l) trunk_col_pack(f.conv_w[l].data(), C, C, C, wc.data() + l * 9 * C * C);
                TC_TRY(upload(m.col_w_stem, ws));
                TC_TRY(upload(m.col_w_conv, wc));
                TC_TRY(upload(m.col_bias, ball));
                TC_TRY(make_map(&m.map_col_stem, m.col_w_stem.p, (uint64_t)9 * C, 16, (uint32_t)(3 * C), 16));
                TC_TRY(make_map(&m.map_col_conv, m.col_w_conv.p, std::max<size_t>(1, L) * 9 * C, C, (uint32_t)(3 * C), (uint32_t)C));
            }
        }
    }
    const size_t budget = 200 * 1024;
    m.conv_stages = pick_stages(m.block_k, C, C <= 64 ? 48 * 1024 : budget);
    m.fc1_stages = pick_stages(64, m.fc1_block_n, m.fc1_block_n > 64 ? 200 * 1024 : std::min<size_t>(96 * 1024, m.side_smem_budget));
    if (const char *fs = getenv("TACULT_FC1_STAGES"))
        if (atoi(fs) >= 2 && atoi(fs) <= MAX_STAGES) m.fc1_stages = atoi(fs);
    TC_CUDA(cudaFuncSetAttribute(k_gemm_tc<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    TC_CUDA(cudaFuncSetAttribute(k_gemm_tc<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    TC_CUDA(cudaFuncSetAttribute(k_gemm_tc<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    TC_CUDA(cudaFuncSetAttribute(k_gemm_tc<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    return TC_OK;
}

int NetBF16::upload(const FoldedNet &f, int capacity, bool share_sm)
{
    loaded = false;
    const int C = f.d.channels, E = f.d.embedding_size;
    if (!(C == 16 || (C % 32 == 0 && C <= 256))) { why_not = "bf16 tensor path needs channels == 16 or a multiple of 32 up to 256"; return TC_OK; }
    if (E % 16 != 0) { why_not = "bf16 tensor path needs embedding_size to be a multiple of 16"; return TC_OK; }
    if (((size_t)81 * C * 2) % 16 != 0) { why_not = "fc1 row pitch must be a multiple of 16 bytes"; return TC_OK; }
    delete impl;
    impl = new NetBF16Impl();
    int rc = upload_impl(*impl, f, capacity, share_sm);
    if (rc != TC_OK) {
        why_not = tc_last_error();
        delete impl;
        impl = nullptr;
        return rc;
    }
    loaded = true;
    why_not.clear();
    return TC_OK;
}

static void launch_gemm(const NetBF16Impl &m, int block_k, const CUtensorMap &a, const CUtensorMap &b, const GemmParams &P,
                        int max_tiles, cudaStream_t st)
{
    const size_t smem = gemm_smem_bytes(block_k, P.block_n, P.stages, P.variant);
    int tmem_cols = 32;
    while (tmem_cols < 2 * P.block_n) tmem_cols <<= 1;
    int per_sm = (int)std::min<size_t>(4, (220 * 1024) / smem);
    per_sm = std::min(per_sm, 512 / tmem_cols);      // every resident CTA must get its TMEM columns
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(max_tiles, m.sm_count * per_sm);
    if (grid < 1) grid = 1;
    if (P.mode == OUT_HEADS) launch_chain(k_gemm_tc<64, true>, grid, GEMM_THREADS, smem, st, a, b, P);
    else if (block_k == 64) launch_chain(k_gemm_tc<64, false>, grid, GEMM_THREADS, smem, st, a, b, P);
    else if (block_k == 32) launch_chain(k_gemm_tc<32, false>, grid, GEMM_THREADS, smem, st, a, b, P);
    else launch_chain(k_gemm_tc<16, false>, grid, GEMM_THREADS, smem, st, a, b, P);
}

void launch_heads_f32(int batch, const int *count_dev, int E, const float *emb, const float *W, const float *bias,
                      float *pi, float *v, float *logits, cudaStream_t st);

// runs the stem and the first n_convs trunk convolutions; returns the index of the buffer holding the result
// (or -1 when the result is in `flat`).  n_convs == all: the last conv writes the compact fc1 input.
static int run_trunk(NetBF16Impl &m, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                     int n_convs, bool to_flat, cudaStream_t st, Profiler *prof, int &launches)
{
    const int C = m.d.channels, I = m.d.in_planes;
    const int total = 2 * m.d.num_residual_blocks;
    if (m.fused && to_flat && n_convs >= total) {
        prof->begin(TC_K_NN_CONV, st);
        if (m.col)
            launch_trunk_col(m.map_col_stem, m.map_col_conv, m.col_bias.p, total, I, batch, count_dev, states, obs, m.flat.p,
                             m.sm_count, st);
        else
            launch_trunk_fused(C, m.map_stem_w, m.map_conv_w_all, m.map_bias_tiles, m.fused_G, m.fused_T, total, I, batch,
                               count_dev, states, obs, m.flat.p, m.sm_count, st);
        prof->end(st);
        ++launches;
        return -1;
    }
    const bool stem_flat = to_flat && n_convs == 0;
    prof->begin(TC_K_NN_STEM, st);
    if (m.stem_on_tensor) {
        if (m.stem_k == 16) k_encode_planes_bf16<16><<<ceil_div(batch, 8), 256, 0, st>>>(batch, count_dev, states, obs, I, m.planes.p);
        else k_encode_planes_bf16<32><<<ceil_div(batch, 8), 256, 0, st>>>(batch, count_dev, states, obs, I, m.planes.p);
        GemmParams S{};
        S.count_dev = count_dev;
        S.batch = batch;
        S.rows_per_unit = ROWS_PER_POS;
        S.taps = 9;
        S.k_chunks = 1;
        S.block_n = C;
        S.n_blocks = 1;
        S.a_row0 = PAD0;
        S.b_tap_rows = C;
        S.variant = m.conv_variant;
        S.stages = pick_stages(m.stem_k, C, 48 * 1024, S.variant);
        S.mode = stem_flat ? OUT_COMPACT_BF16 : OUT_PADDED_BF16;
        S.relu = 1;
        S.ldc = C;
        S.bias = m.stem_b.p;
        S.residual = nullptr;
        S.out = stem_flat ? (void *)m.flat.p : (void *)m.act[0].p;
        launch_gemm(m, m.stem_k, S.variant ? m.map_planes_sh : m.map_planes, m.map_stem_w, S, (batch * ROWS_PER_POS + 127) / 128, st);
        launches += 2;
    } else {
        k_stem_bf16<<<batch, 256, sizeof(float) * I * 121, st>>>(batch, count_dev, states, obs, I, C, m.stem_w.p,
                                                                 m.stem_b.p, stem_flat ? m.flat.p : m.act[0].p,
                                                                 stem_flat ? 1 : 0);
        ++launches;
    }
    prof->end(st);
    if (stem_flat) return -1;
    int x = 0, y = 1, z = 2;
    const int max_tiles = ((batch * ROWS_PER_POS + 127) / 128);
    int cur = x;
    for (int l = 0; l < n_convs && l < total; ++l) {
        const bool second = (l & 1) != 0;
        const bool last = to_flat && l == n_convs - 1;
        GemmParams P{};
        P.count_dev = count_dev;
        P.batch = batch;
        P.rows_per_unit = ROWS_PER_POS;
        P.taps = 9;
        P.k_chunks = C / m.block_k;
        P.block_n = C;
        P.n_blocks = 1;
        P.a_row0 = PAD0;
        P.b_tap_rows = C;
        P.variant = m.conv_variant;
        P.stages = P.variant ? pick_stages(m.block_k, C, 64 * 1024, P.variant) : m.conv_stages;
        P.mode = last ? OUT_COMPACT_BF16 : OUT_PADDED_BF16;
        P.relu = 1;
        P.ldc = C;
        P.bias = m.conv_b[l].p;
        const int src = second ? y : x;
        const int dst = second ? z : y;
        P.residual = second ? m.act[x].p : nullptr;
        P.out = last ? (void *)m.flat.p : (void *)m.act[dst].p;
        prof->begin(TC_K_NN_CONV, st);
        launch_gemm(m, m.block_k, P.variant ? m.map_act_sh[src] : m.map_act[src], m.map_conv_w[l], P, max_tiles, st);
        prof->end(st);
        ++launches;
        cur = dst;
        if (second) { int t = x; x = z; z = t; }
        if (last) return -1;
    }
    return cur;
}

int NetBF16::forward(int batch, const int *count_dev, const uint32_t *states, const float *obs, float *pi, float *v,
                     float *logits, cudaStream_t st, Profiler *prof, int masked)
{
    if (!loaded || !impl) {
        set_error("bf16 evaluator: %s", why_not.c_str());
        return TC_ESTATE;
    }
    NetBF16Impl &m = *impl;
    Profiler none;
    if (!prof) prof = &none;
    int launches = 0;
    const int C = m.d.channels, E = m.d.embedding_size;
    run_trunk(m, batch, count_dev, states, obs, 2 * m.d.num_residual_blocks, true, st, prof, launches);
    if (m.fc1_heads && !logits) {
        prof->begin(TC_K_NN_FC1, st);
        TC_TRY(launch_fc1_heads(m.map_flat, m.map_fc1, m.map_emb16, m.map_head_w, batch, count_dev, 81 * C, E,
                                m.fc1_heads_stages, m.fc1_b.p, m.emb16.p, m.tile_ready.p, m.head_b96.p, pi, v,
                                masked ? states : nullptr, masked && states, st));
        prof->end(st);
        return launches + 1;
    }
    if (masked) {
        set_error("the masked soft-max needs the fused fc1+heads kernel (single partition, embedding a multiple of 64, "
                  "TACULT_FC1_HEADS not 0)");
        return TC_ESTATE;
    }
    GemmParams P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.rows_per_unit = 1;
    P.taps = 1;
    P.k_chunks = (81 * C + 63) / 64;
    P.block_n = m.fc1_block_n;
    P.n_blocks = E / m.fc1_block_n;
    P.a_row0 = 0;
    P.b_tap_rows = 0;
    P.stages = m.fc1_stages;
    P.mode = m.heads_on_tensor ? OUT_PLAIN_BF16 : OUT_F32;
    P.relu = 1;
    P.ldc = E;
    P.bias = m.fc1_b.p;
    P.residual = nullptr;
    P.out = m.heads_on_tensor ? (void *)m.emb16.p : (void *)m.emb.p;
    prof->begin(TC_K_NN_FC1, st);
    if (m.fc1_split_k && m.heads_on_tensor && m.fc1_block_n == 64)
        TC_TRY(launch_fc1_splitk(m.map_flat, m.map_fc1, batch, count_dev, 81 * C, E, m.fc1_stages, m.fc1_b.p, m.emb16.p, st));
    else
        launch_gemm(m, 64, m.map_flat, m.map_fc1, P, ((batch + 127) / 128) * P.n_blocks, st);
    prof->end(st);
    prof->begin(TC_K_NN_HEADS, st);
    if (m.heads_on_tensor) {
        GemmParams H{};
        H.count_dev = count_dev;
        H.batch = batch;
        H.rows_per_unit = 1;
        H.taps = 1;
        H.k_chunks = E / 64;
        H.block_n = 96;
        H.n_blocks = 1;
        H.a_row0 = 0;
        H.b_tap_rows = 0;
        H.stages = pick_stages(64, 96, std::min<size_t>(96 * 1024, m.side_smem_budget));
        H.mode = OUT_HEADS;
        H.ldc = 81;
        H.bias = m.head_b96.p;
        H.out = pi;
        H.out_v = v;
        H.out_logits = logits;
        launch_gemm(m, 64, m.map_emb16, m.map_head_w, H, (batch + 127) / 128, st);
    } else {
        launch_heads_f32(batch, count_dev, E, m.emb.p, m.head_w.p, m.head_b.p, pi, v, logits, st);
    }
    prof->end(st);
    launches += 2;
    return launches;
}

// test hook (tc_debug_trunk): activations after the stem and n_convs convolutions, as f32 [batch][81][C]
int debug_trunk_bf16(NetBF16 &net, int batch, const uint32_t *states_dev, int n_convs, float *out_dev, cudaStream_t st)
{
    if (!net.loaded || !net.impl) {
        set_error("bf16 evaluator: %s", net.why_not.c_str());
        return TC_ESTATE;
    }
    Profiler none;
    int launches = 0;
    int buf = run_trunk(*net.impl, batch, nullptr, states_dev, nullptr, n_convs, false, st, &none, launches);
    k_unpad_to_f32<<<256, 256, 0, st>>>(batch, net.impl->d.channels, net.impl->act[buf].p, 0, out_dev);
    return TC_OK;
}

}  // namespace tc
