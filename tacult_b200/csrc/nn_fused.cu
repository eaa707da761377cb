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
// Roles (320 threads): warp 0 streams the next layer's weights with TMA while the current layer computes;
// warp 1 issues tcgen05.mma (9 taps x C/16 per tile) and commits one mbarrier per tile; warps 2-9 are two
// epilogue groups (tcgen05.ld -> bias (+ residual from smem) -> ReLU -> bf16 -> swizzled st.shared) and also
// build the observation planes from the packed positions at the start of a group.
#include <algorithm>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "tcgen05.cuh"

namespace tc {

constexpr int FUSED_THREADS = 320;
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
    const float *bias;       // [(1 + n_conv)][C]
    __nv_bfloat16 *out;      // compact [row][81][C] (fc1 input)
};

template <int SW>
__device__ __forceinline__ uint32_t swz(uint32_t off)       // byte offset inside a 1024-aligned buffer -> swizzled
{
    return off ^ (((off >> 7) & (uint32_t)(SW / 16 - 1)) << 4);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int C>
__global__ void __launch_bounds__(FUSED_THREADS, 1) k_trunk_fused(const __grid_constant__ CUtensorMap tmW_stem,
                                                                  const __grid_constant__ CUtensorMap tmW_conv,
                                                                  FusedParams P)
{
    constexpr int SW = 2 * C;                                   // activation row = one swizzle span
    constexpr int W_TAP = (C * SW + 1023) & ~1023;              // one conv tap [C x C]
    constexpr int WS_TAP = (C * 32 + 1023) & ~1023;             // one stem tap [C x 16]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int rows = LEAD + 128 * P.T + 16;
    const int buf_bytes = (rows * SW + 1023) & ~1023;
    uint8_t *X = smem, *Y = smem + buf_bytes, *W = smem + 2 * buf_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(W + 9 * W_TAP);
    uint64_t *tdone = bars, *wfull = bars + MAX_TILES, *wfree = bars + MAX_TILES + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + MAX_TILES + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    const int n_groups = (count + P.G - 1) / P.G;
    const int n_layers = 1 + P.n_conv;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < P.T * C) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int t = 0; t < MAX_TILES; ++t) mbar_init(tdone + t, 1);
        mbar_init(wfull, 1);
        mbar_init(wfree, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmW_stem);
        tma_prefetch_desc(&tmW_conv);
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    // pad rows of both activation buffers must read as zero for ever: clear everything once
    for (int i = threadIdx.x; i < 2 * buf_bytes / 16; i += FUSED_THREADS)
        reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    uint32_t lcount = 0;       // layers executed so far by this CTA (mbarrier phase = lcount & 1)
    if (warp == 0 && lane == 0 && (int)blockIdx.x < n_groups) {
        mbar_expect_tx(wfull, 9u * C * 32u);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(W + tap * WS_TAP, &tmW_stem, wfull, 0, tap * C);
    }

    for (int group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int pos0 = group * P.G;
        // ---- observation planes of the group -> channels 0..15 of the group's rows in Y (activation layout:
        //      the stem's UMMAs read K = 16 of it; pad rows are never touched and stay zero) ----
        if (warp >= 2) {
            for (int idx = threadIdx.x - 64; idx < P.G * 81; idx += FUSED_THREADS - 64) {
                const int g = idx / 81, a = idx - 81 * g;
                const int row = pos0 + g;
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = 0.f;
                if (row < count) {
                    if (P.states) {
                        uint32_t wd[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) wd[k] = P.states[8 * (size_t)row + k];
                        Pos p = unpack(wd);
                        uint32_t lw[3];
                        legal_words(p, summarize(p), lw);
                        const bool mine = action_bit(p.mine, a), occ = action_bit(p.occ, a);
                        v[0] = mine ? 1.f : 0.f;
                        v[1] = (occ && !mine) ? 1.f : 0.f;
                        v[2] = action_bit(lw, a) ? 1.f : 0.f;
                        v[3] = 1.f;
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c < P.in_planes) v[c] = P.obs[((size_t)row * P.in_planes + c) * 81 + a];
                    }
                }
                uint4 o[2];
                __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(o);
#pragma unroll
                for (int c = 0; c < 8; ++c) op[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
                const uint32_t off = (uint32_t)(LEAD + 100 * g + (a / 9) * 10 + a % 9) * (uint32_t)SW;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off)) = o[0];
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off + 16)) = o[1];
            }
            fence_async_smem();
        }
        __syncthreads();

        for (int l = 0; l < n_layers; ++l, ++lcount) {
            const uint32_t ph = lcount & 1u;
            // layer 0: planes(Y) -> X;  odd l (first conv of a block): X -> Y;  even l > 0: Y -> X, + residual X
            uint8_t *src = (l == 0 || (l & 1) == 0) ? Y : X;
            uint8_t *dst = (l & 1) ? Y : X;
            const bool last = l == n_layers - 1;
            if (warp == 0) {
                if (lane == 0) {
                    // once this layer's MMAs have consumed the weights, stream in the next layer's
                    mbar_wait(wfree, ph);
                    const bool more_layers = l + 1 < n_layers;
                    const bool more_groups = group + (int)gridDim.x < n_groups;
                    if (more_layers) {
                        mbar_expect_tx(wfull, 9u * C * (uint32_t)SW);
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_2d(W + tap * W_TAP, &tmW_conv, wfull, 0, (l * 9 + tap) * C);
                    } else if (more_groups) {
                        mbar_expect_tx(wfull, 9u * C * 32u);
                        for (int tap = 0; tap < 9; ++tap) tma_load_2d(W + tap * WS_TAP, &tmW_stem, wfull, 0, tap * C);
                    }
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    const uint32_t idesc = make_instr_desc(C);
                    mbar_wait(wfull, ph);
                    tc_fence_after();
                    const uint32_t s_addr = smem_u32(src), w_addr = smem_u32(W);
                    // Consecutive UMMAs into the SAME accumulator serialise on the tensor pipe's latency (N = C is a
                    // few cycles of work each), so the issue order interleaves the tiles of a chunk: tap-major,
                    // tile-minor.  Two chunks per layer let the epilogue of the first overlap the MMAs of the second.
                    const int half = (P.T + 1) / 2;
                    for (int t0 = 0; t0 < P.T; t0 += half) {
                        const int t1 = min(P.T, t0 + half);
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const int shift0 = LEAD + (tap / 3 - 1) * 10 + (tap % 3 - 1);
                            if (l == 0) {
                                const uint64_t bdesc = make_smem_desc<32>(w_addr + (uint32_t)(tap * WS_TAP));
                                for (int t = t0; t < t1; ++t)
                                    umma_bf16(tmem_base + (uint32_t)(t * C),
                                              make_smem_desc<SW>(s_addr + (uint32_t)((shift0 + 128 * t) * SW)), bdesc, idesc,
                                              (uint32_t)(tap != 0));
                            } else {
#pragma unroll
                                for (int k = 0; k < C / 16; ++k) {
                                    const uint64_t bdesc = make_smem_desc<SW>(w_addr + (uint32_t)(tap * W_TAP) + 32 * k);
                                    for (int t = t0; t < t1; ++t)
                                        umma_bf16(tmem_base + (uint32_t)(t * C),
                                                  make_smem_desc<SW>(s_addr + (uint32_t)((shift0 + 128 * t) * SW) + 32 * k), bdesc,
                                                  idesc, (uint32_t)((tap | k) != 0));
                                }
                            }
                        }
                        for (int t = t0; t < t1; ++t) umma_commit(tdone + t);
                    }
                    umma_commit(wfree);
                }
            } else {
                // ===== epilogue: two groups of four warps, tiles alternate between the groups =====
                const int ew = warp - 2, quarter = warp & 3, egroup = ew >> 2;
                float breg[C];
                const float *bias = P.bias + (size_t)l * C;
#pragma unroll
                for (int c = 0; c < C; ++c) breg[c] = __ldg(bias + c);
                const bool residual = l > 0 && (l & 1) == 0;
                for (int t = egroup; t < P.T; t += 2) {
                    const int q = 128 * t + 32 * quarter + lane;          // padded row inside the group
                    const int g = q / 100, rem = q - 100 * g;
                    const int r = rem / 10, cc = rem - 10 * r;
                    const int row = pos0 + g;
                    const bool valid = g < P.G && row < count && rem < 90 && cc != 9;
                    mbar_wait(tdone + t, ph);
                    tc_fence_after();
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
                                float v[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(acc[j + k]) + breg[c0 + j + k];
                                const uint32_t so = swz<SW>(row_off + (uint32_t)(c0 + j) * 2u);
                                if (residual) {
                                    uint4 rv = *reinterpret_cast<const uint4 *>(X + so);
                                    const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        float2 f = __bfloat1622float2(rp[k]);
                                        v[2 * k] += f.x;
                                        v[2 * k + 1] += f.y;
                                    }
                                }
                                uint4 ov;
                                __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    op[k] = __floats2bfloat162_rn(fmaxf(v[2 * k], 0.f), fmaxf(v[2 * k + 1], 0.f));
                                if (last)
                                    *reinterpret_cast<uint4 *>(P.out + ((size_t)row * 81 + r * 9 + cc) * C + c0 + j) = ov;
                                else
                                    *reinterpret_cast<uint4 *>(dst + so) = ov;
                            }
                        }
                    }
                }
                fence_async_smem();       // activations written through the generic proxy feed the next layer's UMMAs
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static size_t fused_smem_bytes(int C, int T)
{
    const size_t SW = 2 * C;
    const size_t rows = LEAD + 128 * T + 16;
    const size_t buf = (rows * SW + 1023) & ~(size_t)1023;
    const size_t wtap = (C * SW + 1023) & ~(size_t)1023;
    return 1024 + 2 * buf + 9 * wtap + (MAX_TILES + 2) * 8 + 16;
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

int launch_trunk_fused(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, int G, int T, int n_conv,
                       int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                       const float *bias_all, __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st)
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
    P.bias = bias_all;
    P.out = out_flat;
    const size_t smem = fused_smem_bytes(C, T);
    const int groups = (batch + G - 1) / G;
    const int grid = std::max(1, std::min(groups, sm_count));
    if (C == 16) {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_trunk_fused<16><<<grid, FUSED_THREADS, smem, st>>>(w_stem, w_conv, P);
    } else if (C == 32) {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_trunk_fused<32><<<grid, FUSED_THREADS, smem, st>>>(w_stem, w_conv, P);
    } else {
        TC_CUDA(cudaFuncSetAttribute(k_trunk_fused<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_trunk_fused<64><<<grid, FUSED_THREADS, smem, st>>>(w_stem, w_conv, P);
    }
    return TC_OK;
}

}  // namespace tc
