// Fused trunk for narrow nets (C = 16 / 32 / 64): conv_in and every residual-block convolution of
// `_UtacNNet` (/root/reference/src/tacult/network.py:84-89, ResidualBlock 35-41) in ONE kernel, activations
// resident in shared memory from the packed bitboards to the fc1 input.
//
// Why: at C = 32 a 3x3 conv layer is 144 flop per byte of activation traffic — below the B200 ridge — so the
// layer-per-kernel path is bound by L2/HBM round trips, not by the tensor pipe.  Here a CTA owns G positions
// (G*100 padded rows, T = ceil(G*100/128) UMMA tiles): X/Y activation buffers (bf16, rows of 2C bytes in the
// TMA/UMMA swizzle pattern) and the layer's 9 weight taps live in shared memory, the T accumulators
// (T*C <= 512 columns) in TMEM.  A 3x3 tap is the same smem tile read through a UMMA descriptor whose start
// address is moved by 10*dy+dx rows — measured exact on B200 with matrix base offset 0, i.e. the swizzle XOR is
// a function of the absolute shared-memory address.
//
// Everything linear is done by the tensor pipe, so the epilogue is only ReLU + bf16 pack + store:
//   bias      = one K=16 UMMA  (all-ones A tile) x (bias split in bf16 hi+lo on two K slots)
//   residual  = C/16 UMMAs     (block input X, unshifted) x (identity tile)
// (adding the bias in the epilogue instead was measured 12 % slower, presumably because its shared-memory reads compete with the UMMA
// operand fetches, and the L1TEX/smem data path is this kernel's busiest unit - 76 % in ncu.)
// A UMMA with N = C is only a few cycles of tensor work but costs ~45-55 cycles end to end, so one issuing thread
// cannot keep the pipe busy: four issuer warps (one elected lane each, descriptors in uniform registers) issue the
// UMMAs of tiles t = i (mod 4), eight epilogue warps (two per TMEM lane quarter) drain the accumulators, and warp 0
// streams the NEXT layer's weights into the other half of a double buffer with TMA.  Layers are not separated by a
// block-wide barrier: tile t of layer l+1 starts as soon as the epilogues of tiles t-1, t, t+1 of layer l have arrived
// on per-tile mbarriers (two banks, see below).  The prologue overlaps the previous kernel's tail (griddep_wait).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "rules_warp.cuh"
#include "tcgen05.cuh"

namespace tc {

constexpr int ISSUERS = 4;               // warps 1..4: one elected lane each issues the UMMAs of tiles t = i (mod 4)
constexpr int EPI_WARPS = 8;             // warps 5..12: two per TMEM lane quarter
constexpr int FUSED_THREADS = 32 * (1 + ISSUERS + EPI_WARPS);
constexpr int LEAD = 16;                 // zero rows in front of the first position of a group
constexpr int MAX_TILES = 16;

struct FusedParams {
    const int *count_dev;
    int batch;
    int G;                   // positions per group
    int T;                   // 128-row tiles per group
    int n_conv;              // 2 * residual blocks
    int in_planes;
    const uint32_t *states;  // packed positions, or
    const float *obs;        // dense planes [row][in_planes][81]
    __nv_bfloat16 *out;      // compact [row][81][C] (fc1 input)
    long long *timeline;     // optional: clock64 stamps of CTA 0 (debugging / profiling aid), else null
    int pdl_trigger;         // griddepcontrol.launch_dependents at kernel entry (common.cuh)
};

template <int SW>
__device__ __forceinline__ uint32_t swz(uint32_t off)       // byte offset inside a 1024-aligned buffer -> swizzled
{
    return off ^ (((off >> 7) & (uint32_t)(SW / 16 - 1)) << 4);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint32_t relu_pack(uint32_t a_bits, uint32_t b_bits, __nv_bfloat162 zero2)
{
    __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(__uint_as_float(a_bits), __uint_as_float(b_bits)), zero2);
    return *reinterpret_cast<uint32_t *>(&h);
}

// descriptor = constant high word per layout + (shared address >> 4) in the low word
template <int SW>
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr)
{
    constexpr uint32_t layout = SW == 128 ? 2u : (SW == 64 ? 4u : 6u);
    constexpr uint32_t hi = ((8u * SW) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | (uint64_t)(saddr >> 4);
}

template <int C>
__global__ void __launch_bounds__(FUSED_THREADS, 1) k_trunk_fused(const __grid_constant__ CUtensorMap tmW_stem,
                                                                  const __grid_constant__ CUtensorMap tmW_conv,
                                                                  const __grid_constant__ CUtensorMap tmBias,
                                                                  FusedParams P)
{
    constexpr int SW = 2 * C;                                   // activation row = one swizzle span
    constexpr int W_TAP = (C * SW + 1023) & ~1023;              // one conv tap [C x C]
    constexpr int WS_TAP = (C * 32 + 1023) & ~1023;             // one stem tap / bias tile [C x 16]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int rows = LEAD + 128 * P.T + 16;
    const int buf_bytes = (rows * SW + 1023) & ~1023;
    uint8_t *X = smem, *Y = smem + buf_bytes;
    constexpr int W_BUF = 9 * W_TAP + WS_TAP;                   // 9 taps + the bias tile of one layer
    uint8_t *W = smem + 2 * buf_bytes;                          // two layers' weights: layer n lives in buffer n & 1
    uint8_t *Id = W + 2 * W_BUF;                                // identity [C x C]
    uint8_t *Ones = Id + W_TAP;                                 // 128 rows x 16 ones
    uint64_t *bars = reinterpret_cast<uint64_t *>(Ones + 4096);
    uint64_t *tdone = bars;                                     // [t]: the UMMAs of tile t have completed
    // [layer & 1][t]: the epilogue of tile t has written its activations.  Two banks, because a parity wait must never
    // fall two phases behind: a neighbouring tile may run a whole layer ahead of the issuer that waits on it, not two.
    uint64_t *aready = bars + MAX_TILES;
    uint64_t *wfull = bars + 3 * MAX_TILES, *wfree = bars + 3 * MAX_TILES + 2;      // [2] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * MAX_TILES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + P.n_conv;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < P.T * C) tmem_cols <<= 1;

    // ---- prologue that depends on nothing the previous kernel wrote: it overlaps that kernel's tail (programmatic
    //      dependent launch, common.cuh) ----
    griddep_launch(P.pdl_trigger == 1);
    // optional marks of CTA 0 (TACULT_FUSED_TIMELINE): kernel entry, prologue done, dependency wait over, first planes
    // encoded, every group done, exit — where a launch spends the time that is not layer work
    long long *marks = (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) ? P.timeline + 8 * 32 : nullptr;
    if (marks) marks[0] = clock64();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW_stem);
        tma_prefetch_desc(&tmW_conv);
        tma_prefetch_desc(&tmBias);
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    // pad rows of both activation buffers must read as zero for ever: clear everything once
    for (int i = threadIdx.x; i < 2 * buf_bytes / 16; i += FUSED_THREADS)
        reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < W_TAP / 16; i += FUSED_THREADS) reinterpret_cast<uint4 *>(Id)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 4096 / 16; i += FUSED_THREADS)
        reinterpret_cast<uint4 *>(Ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);

    // ---- from here on the leaf count and the leaf positions written by the descent kernel are read ----
    if (marks) marks[1] = clock64();
    griddep_wait();
    if (marks) marks[2] = clock64();
    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    // The live positions are dealt out evenly: CTA b owns positions [b * per, (b + 1) * per) and walks them in groups of
    // P.G (the largest group the shared-memory carve-up holds) plus one smaller remainder group, which costs fewer 128-row
    // tiles than equal groups (21 positions: 14 + 7 -> 11 + 6 tiles, against 11 + 11 -> 9 + 9).
    const int per = (count + (int)gridDim.x - 1) / (int)gridDim.x;
    const int p_begin = min(count, (int)blockIdx.x * per), p_end = min(count, p_begin + per);
    const int n_my_groups = (p_end - p_begin + P.G - 1) / P.G;
    if (n_my_groups == 0 && threadIdx.x == 0) griddep_launch(P.pdl_trigger == 2);
    if (threadIdx.x == 0) {
        for (int t = 0; t < MAX_TILES; ++t) mbar_init(tdone + t, 1);
        for (int t = 0; t < 2 * MAX_TILES; ++t) mbar_init(aready + t, 4);
        for (int b = 0; b < 2; ++b) { mbar_init(wfull + b, 1); mbar_init(wfree + b, (uint32_t)ISSUERS); }
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < C)            // identity, K-major rows of SW bytes in the swizzle pattern
        *reinterpret_cast<__nv_bfloat16 *>(Id + swz<SW>((uint32_t)threadIdx.x * SW + (uint32_t)threadIdx.x * 2u)) =
            __float2bfloat16_rn(1.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);

    uint32_t lcount = 0;       // layers executed so far by this CTA (mbarrier phase = lcount & 1)
    if (warp == 0 && lane == 0 && n_my_groups > 0) {
        mbar_expect_tx(wfull, 10u * C * 32u);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(W + tap * WS_TAP, &tmW_stem, wfull, 0, tap * C);
        tma_load_2d(W + 9 * W_TAP, &tmBias, wfull, 0, 0);
    }

    const uint32_t x_addr = smem_u32(X), y_addr = smem_u32(Y);
    const uint64_t ones_desc = desc_of<32>(smem_u32(Ones));
    const uint32_t id_addr = smem_u32(Id);
    const uint32_t idesc = make_instr_desc(C);
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);

    for (int group = 0; group < n_my_groups; ++group) {
        const int pos0 = p_begin + group * P.G;
        const int G = min(P.G, p_end - pos0);
        const int T = (100 * G + 127) / 128;
        // ---- observation planes of the group -> channels 0..15 of the group's rows in Y (activation layout:
        //      the stem's UMMAs read K = 16 of it; pad rows are never touched and stay zero) ----
        if (warp > ISSUERS && P.states) {
            // one warp per position: the legal mask is computed ONCE, by the warp (rules_warp.cuh), then lane writes the
            // cells lane, lane + 32, lane + 64 (the per-cell scalar rules this replaced took 8 k cycles per group)
            for (int g = warp - 1 - ISSUERS; g < G; g += EPI_WARPS) {
                const uint32_t w = lane < 8 ? P.states[8 * (size_t)(pos0 + g) + lane] : 0u;
                uint32_t wd[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) wd[k] = __shfl_sync(0xFFFFFFFFu, w, k);
                const Pos p = unpack(wd);
                uint32_t am[3], mm[3], oo[3];
                int status;
                warp_legal_mask<true>(p, lane, am, status, mm, oo);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int a = lane + 32 * q;
                    if (a < 81) {
                        const bool mine = (mm[q] >> lane) & 1u, occ = (oo[q] >> lane) & 1u, legal = (am[q] >> lane) & 1u;
                        const int r = (a * 57) >> 9, cc = a - 9 * r;
                        const uint32_t off = (uint32_t)(LEAD + 100 * g + 10 * r + cc) * (uint32_t)SW;
                        // bf16 1.0 = 0x3F80: planes mine, theirs | legal, ones; channels 4..15 are zero
                        *reinterpret_cast<uint4 *>(Y + swz<SW>(off)) =
                            make_uint4((mine ? 0x3F80u : 0u) | ((occ && !mine) ? 0x3F800000u : 0u), (legal ? 0x3F80u : 0u) | 0x3F800000u, 0u, 0u);
                        *reinterpret_cast<uint4 *>(Y + swz<SW>(off + 16)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
            fence_async_smem();
        } else if (warp > ISSUERS) {
            for (int idx = threadIdx.x - 32 * (1 + ISSUERS); idx < G * 81; idx += 32 * EPI_WARPS) {
                const int g = idx / 81, a = idx - 81 * g;
                const int row = pos0 + g;
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = (row < count && c < P.in_planes) ? P.obs[((size_t)row * P.in_planes + c) * 81 + a] : 0.f;
                uint4 o0, o1;
                o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                const uint32_t off = (uint32_t)(LEAD + 100 * g + (a / 9) * 10 + a % 9) * (uint32_t)SW;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off)) = o0;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off + 16)) = o1;
            }
            fence_async_smem();
        }
        __syncthreads();
        if (marks && group < 4) marks[3 + 2 * group] = clock64();

        // Layers are NOT separated by a block-wide barrier: tile t of layer l+1 only needs the activations of tiles
        // t-1, t, t+1 of layer l (taps reach 11 rows), so its UMMAs start as soon as those three epilogues have
        // arrived on aready[], while the tail of layer l is still draining.  Weights are double buffered.
        for (int l = 0; l < n_layers; ++l, ++lcount) {
            const uint32_t ph = lcount & 1u;                    // tdone / aready complete once per layer
            const uint32_t wb = lcount & 1u, wph = (lcount >> 1) & 1u;      // weight buffer of this layer and its phase
            const bool stamp = P.timeline && blockIdx.x == 0 && group == 0 && lane == 0;
            // layer 0: planes(Y) -> X;  odd l (first conv of a block): X -> Y;  even l > 0: Y -> X, + residual X
            const bool to_y = (l & 1) != 0;
            const uint32_t s_addr = to_y ? x_addr : y_addr;
            uint8_t *dst = to_y ? Y : X;
            const bool last = l == n_layers - 1;
            const bool residual = l > 0 && !to_y;
            // late trigger (TACULT_PDL_TRIGGER=2): once every CTA is in the last layer of its last group, fc1's CTAs may
            // become resident on the SMs that free up first and run their prologue behind this kernel's tail
            if (last && group == n_my_groups - 1 && threadIdx.x == 0) griddep_launch(P.pdl_trigger == 2);
            if (warp == 0) {
                if (lane == 0) {
                    // stream the NEXT layer's weights into the other buffer once the layer before this one (its
                    // previous user) has been consumed by the tensor pipe
                    const bool more_layers = l + 1 < n_layers;
                    const bool more_groups = group + 1 < n_my_groups;
                    if (more_layers || more_groups) {
                        const uint32_t nb = wb ^ 1u;
                        uint8_t *Wn = W + nb * W_BUF;
                        if (lcount >= 1) mbar_wait(wfree + nb, ((lcount - 1) >> 1) & 1u);
                        if (more_layers) {
                            mbar_expect_tx(wfull + nb, 9u * C * (uint32_t)SW + C * 32u);
                            for (int tap = 0; tap < 9; ++tap)
                                tma_load_2d(Wn + tap * W_TAP, &tmW_conv, wfull + nb, 0, (l * 9 + tap) * C);
                            tma_load_2d(Wn + 9 * W_TAP, &tmBias, wfull + nb, 0, (l + 1) * C);
                        } else {
                            mbar_expect_tx(wfull + nb, 10u * C * 32u);
                            for (int tap = 0; tap < 9; ++tap)
                                tma_load_2d(Wn + tap * WS_TAP, &tmW_stem, wfull + nb, 0, tap * C);
                            tma_load_2d(Wn + 9 * W_TAP, &tmBias, wfull + nb, 0, 0);
                        }
                    }
                }
            } else if (warp <= ISSUERS) {
                const int e = warp - 1;
                // ===== issue the UMMAs of this warp's tiles (warp-uniform code, one elected lane issues); an issuer without
                //       a tile in this group still waits for the layer's weights and releases them, so that the arrivals
                //       of two layers can never mix on a wfree barrier =====
                {
                    if (stamp && warp == 1) P.timeline[8 * l + 0] = clock64();
                    mbar_wait(wfull + wb, wph);
                    tc_fence_after();
                    if (stamp && warp == 1) P.timeline[8 * l + 1] = clock64();
                    if (elect_one()) {
                        const uint32_t w_addr = smem_u32(W + wb * W_BUF);
                        const uint64_t bias_desc = desc_of<32>(w_addr + 9 * W_TAP);
                        // tile by tile, so that a finished tile's epilogue overlaps the UMMAs of the next ones; the four
                        // issuers keep four independent accumulators in flight on the tensor pipe
                        for (int t = e; t < T; t += ISSUERS) {
                            if (l > 0) {
                                // inputs of this tile: the previous layer's rows of tiles t-1, t, t+1 (and the
                                // accumulator columns of tile t, drained by the same epilogue)
                                uint64_t *prev = aready + ((lcount - 1) & 1u) * MAX_TILES;
                                const uint32_t pph = ((lcount - 1) >> 1) & 1u;
                                if (t > 0) mbar_wait(prev + t - 1, pph);
                                mbar_wait(prev + t, pph);
                                if (t + 1 < T) mbar_wait(prev + t + 1, pph);
                                tc_fence_after();
                            }
                            const uint32_t d_tmem = tmem_base + (uint32_t)(t * C);
                            const uint32_t a_tile = s_addr + (uint32_t)((LEAD + 128 * t) * SW);
                            umma_bf16(d_tmem, ones_desc, bias_desc, idesc, 0u);      // accumulator starts as the bias
#pragma unroll 1
                            for (int tap = 0; tap < 9; ++tap) {
                                const int shift = ((tap / 3 - 1) * 10 + (tap % 3 - 1)) * SW;
                                if (l == 0) {
                                    umma_bf16(d_tmem, desc_of<SW>(a_tile + (uint32_t)shift),
                                              desc_of<32>(w_addr + (uint32_t)(tap * WS_TAP)), idesc, 1u);
                                } else {
#pragma unroll
                                    for (int k = 0; k < C / 16; ++k)
                                        umma_bf16(d_tmem, desc_of<SW>(a_tile + (uint32_t)shift + 32 * k),
                                                  desc_of<SW>(w_addr + (uint32_t)(tap * W_TAP) + 32 * k), idesc, 1u);
                                }
                            }
                            if (residual) {
#pragma unroll
                                for (int k = 0; k < C / 16; ++k)
                                    umma_bf16(d_tmem, desc_of<SW>(x_addr + (uint32_t)((LEAD + 128 * t) * SW) + 32 * k),
                                              desc_of<SW>(id_addr + 32 * k), idesc, 1u);
                            }
                            umma_commit(tdone + t);
                        }
                        umma_commit(wfree + wb);
                    }
                    __syncwarp();
                    if (stamp && warp == 1) P.timeline[8 * l + 2] = clock64();
                }
            } else {
                // ===== epilogue: ReLU + bf16 pack + store; two warps per TMEM lane quarter, tiles alternate =====
                const int quarter = warp & 3, egroup = (warp - 1 - ISSUERS) >> 2;
                for (int t = egroup; t < T; t += 2) {
                    const int q = 128 * t + 32 * quarter + lane;          // padded row inside the group
                    const int g = q / 100, rem = q - 100 * g;
                    const int r = rem / 10, cc = rem - 10 * r;
                    const int row = pos0 + g;
                    const bool valid = g < G && row < count && rem < 90 && cc != 9;
                    mbar_wait(tdone + t, ph);
                    tc_fence_after();
                    if (stamp && warp == 1 + ISSUERS && t == egroup) P.timeline[8 * l + 3] = clock64();
                    const uint32_t t_addr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(t * C);
                    const uint32_t row_off = (uint32_t)(LEAD + q) * (uint32_t)SW;
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += (C < 32 ? 16 : 32)) {
                        constexpr int CW = C < 32 ? 16 : 32;
                        uint32_t acc[32];
                        if (CW == 32) tmem_ld32(t_addr + (uint32_t)c0, acc); else tmem_ld16(t_addr + (uint32_t)c0, acc);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < CW; j += 8) {
                                uint4 ov;
                                ov.x = relu_pack(acc[j], acc[j + 1], zero2);
                                ov.y = relu_pack(acc[j + 2], acc[j + 3], zero2);
                                ov.z = relu_pack(acc[j + 4], acc[j + 5], zero2);
                                ov.w = relu_pack(acc[j + 6], acc[j + 7], zero2);
                                if (last)
                                    *reinterpret_cast<uint4 *>(P.out + ((size_t)row * 81 + r * 9 + cc) * C + c0 + j) = ov;
                                else
                                    *reinterpret_cast<uint4 *>(dst + swz<SW>(row_off + (uint32_t)(c0 + j) * 2u)) = ov;
                            }
                        }
                    }
                    // the rows are written through the generic proxy and read by the next layer's UMMAs (async proxy);
                    // the accumulator columns are free again as well
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(aready + wb * MAX_TILES + t);
                }
                if (stamp && warp == 1 + ISSUERS) P.timeline[8 * l + 4] = clock64();
            }
        }
        // group boundary: the next group's planes overwrite Y and its stem overwrites the accumulators
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (marks && group < 4) marks[4 + 2 * group] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
    if (marks) { marks[11] = clock64(); marks[12] = n_my_groups; marks[13] = p_end - p_begin; }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
size_t fused_smem_bytes(int C, int T)
{
    const size_t SW = 2 * C;
    const size_t rows = LEAD + 128 * T + 16;
    const size_t buf = (rows * SW + 1023) & ~(size_t)1023;
    const size_t wtap = (C * SW + 1023) & ~(size_t)1023;
    const size_t wstap = (C * 32 + 1023) & ~(size_t)1023;
    return 1024 + 2 * buf + 2 * (9 * wtap + wstap) + wtap + 4096 + (3 * MAX_TILES + 4) * 8 + 16;
}

bool fused_plan(int C, int &G, int &T)
{
    if (!(C == 16 || C == 32 || C == 64)) return false;
    const int t_tmem = std::min(MAX_TILES, 512 / C);
    T = 0;
    for (int t = t_tmem; t >= 1; --t)
        if (fused_smem_bytes(C, t) <= 227 * 1024) { T = t; break; }
    if (T == 0) return false;
    G = (128 * T) / 100;
    return G >= 1;
}

// bias of every layer as a K-major bf16 tile [C][16]: k=0 holds the bf16 head of the bias, k=1 the remainder
void fused_bias_tiles(const std::vector<float> &bias_all, std::vector<__nv_bfloat16> &out)
{
    const size_t n = bias_all.size();
    out.assign(n * 16, __float2bfloat16_rn(0.f));
    for (size_t i = 0; i < n; ++i) {
        __nv_bfloat16 hi = __float2bfloat16_rn(bias_all[i]);
        __nv_bfloat16 lo = __float2bfloat16_rn(bias_all[i] - __bfloat162float(hi));
        out[i * 16] = hi;
        out[i * 16 + 1] = lo;
    }
}

int launch_trunk_fused(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G,
                       int T, int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states,
                       const float *obs, __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st)
{
    FusedParams P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.G = G;
    P.T = T;
    P.n_conv = n_conv;
    P.in_planes = in_planes;
    P.states = states;
    P.obs = obs;
    P.out = out_flat;
    P.pdl_trigger = pdl_trigger_enabled();
    static long long *timeline_dev = nullptr;
    if (getenv("TACULT_FUSED_TIMELINE")) {
        if (!timeline_dev) { cudaMalloc(&timeline_dev, 8 * 64 * sizeof(long long)); cudaMemset(timeline_dev, 0, 8 * 64 * sizeof(long long)); }
        P.timeline = timeline_dev;
    }
    const size_t smem = fused_smem_bytes(C, T);
    const int grid = std::max(1, std::min(batch, sm_count));     // the kernel sizes its groups from the live count
    if (C == 16) {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_CUDA(launch_chain(k_trunk_fused<16>, grid, FUSED_THREADS, smem, st, w_stem, w_conv, bias_tiles, P));
    } else if (C == 32) {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_CUDA(launch_chain(k_trunk_fused<32>, grid, FUSED_THREADS, smem, st, w_stem, w_conv, bias_tiles, P));
    } else {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_CUDA(launch_chain(k_trunk_fused<64>, grid, FUSED_THREADS, smem, st, w_stem, w_conv, bias_tiles, P));
    }
    if (P.timeline) {
        long long h[8 * 64];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, timeline_dev, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[fused timeline] layer: issuer-0 start | weights ready | issued | first tile done | epilogues done (cycles since layer 0 start)\n");
        for (int l = 0; l <= n_conv; ++l)
            fprintf(stderr, "[fused timeline] %2d: %7lld %7lld %7lld %7lld %7lld\n", l, h[8 * l] - h[0], h[8 * l + 1] - h[0],
                    h[8 * l + 2] - h[0], h[8 * l + 3] - h[0], h[8 * l + 4] - h[0]);
        const long long *m = h + 8 * 32;
        fprintf(stderr, "[fused marks] CTA 0, %lld positions in %lld groups; cycles since kernel entry: prologue done %lld | "
                "dependency wait over %lld | group 0 planes %lld done %lld | group 1 planes %lld done %lld | exit %lld\n",
                m[13], m[12], m[1] - m[0], m[2] - m[0], m[3] - m[0], m[4] - m[0], m[12] > 1 ? m[5] - m[0] : 0,
                m[12] > 1 ? m[6] - m[0] : 0, m[11] - m[0]);
    }
    return TC_OK;
}

}  // namespace tc
