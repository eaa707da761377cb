// Fused trunk, third formulation: dx taps stacked in N (as nn_fused_dx.cu) on a PITCH-16 activation layout, so that the
// +-1-row combine never leaves a warp.
//
// Measured on B200: a UMMA with M = 128, K = 16 costs ~45 cycles whatever N is (16, 32 and 96 measured) - the A fetch
// sets the price - so the trunk's speed is the number of UMMAs.  Stacking the three horizontal taps in N turns 3 UMMAs
// into 1:  acc_dx[m] = sum_dy X[m + P*dy] * W[dy][dx],  out[m] = acc_-1[m-1] + acc_0[m] + acc_+1[m+1].
// With the pitch-10 layout of nn_fused.cu the +-1 neighbours of a row can sit in another warp's TMEM lanes (nn_fused_dx.cu
// pays an exchange through shared memory and loses).  Here every board row owns a 16-row slot
//     row = LEAD + 160*g + 16*r + 1 + c        (slot offsets 0 and 10..15 are zero and never written; slot 9 of a position
//                                               is an all-zero separator row for dy = +-1)
// and a 32-lane TMEM block holds exactly two slots: a cell's left/right neighbours are always in the same warp, the
// combine is two shuffles per value (the SM moves one warp-shuffle per cycle: tools/probes/shfl_rate_probe.cu), and
// tiles need no overlap.  The dy tap is a descriptor shift of +-16 rows = +-1 KB, aligned to the swizzle atom.
// Cost: 160 instead of 100 rows per position, but 7-9 UMMAs per tile and layer instead of 19-21:
// 1.25 * 8 = 10 UMMAs per position and layer against 0.8 * 20 = 16.
//
// STATUS (measured on B200, round 1): exact (tests/test_gpu_net_bf16.py), but SLOWER than k_trunk_fused - opt-in with
// TACULT_FUSED_P16=1: 124 us against 92 us for 4096 positions.  The in-kernel timeline shows why the premise above is
// only half true: a UMMA costs ~40 cycles for the A fetch PLUS ~N/2 cycles for the B fetch (N = 32: ~52, N = 96: ~90),
// i.e. the tensor pipe is fed at ~80-100 B/clk of shared-memory operand bandwidth whatever the shape.  Per position and
// layer this layout moves 10 UMMAs x 7 KB = 70 KB of operands against 15.6 x 5 KB = 78 KB for the pitch-10 kernel - no
// real saving - and small groups (4 positions, T = 5 tiles, bounded by 512 TMEM columns / 96) add per-group overhead.
//
// Structure as k_trunk_fused: 4 issuer warps, 16 epilogue warps (4 groups), 1 TMA warp; accumulators of all T tiles in
// TMEM (T * 3C <= 512 columns); no block-wide barrier between layers (per-tile mbarriers in two banks); weights double
// buffered; bias and residual on the tensor pipe.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "tcgen05.cuh"

namespace tc {

namespace p16 {

constexpr int ISSUERS = 4;
constexpr int EPI_GROUPS = 4;
constexpr int THREADS = 32 * (1 + ISSUERS + 4 * EPI_GROUPS);
constexpr int LEAD = 16;
constexpr int POS_ROWS = 160;            // rows per position: 10 slots of 16
constexpr int MAX_TILES = 10;

struct Params {
    const int *count_dev;
    int batch;
    int G, T;                // positions / 128-row tiles per group
    int n_conv;
    int in_planes;
    const uint32_t *states;
    const float *obs;
    __nv_bfloat16 *out;      // compact [row][81][C]
    long long *timeline;     // optional clock64 stamps of CTA 0's first group (TACULT_FUSED_TIMELINE)
};

template <int SW>
__device__ __forceinline__ uint32_t swz(uint32_t off)
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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t relu_pack(float a, float b, __nv_bfloat162 zero2)
{
    __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(a, b), zero2);
    return *reinterpret_cast<uint32_t *>(&h);
}
template <int SW>
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr)
{
    constexpr uint32_t layout = SW == 128 ? 2u : (SW == 64 ? 4u : 6u);
    constexpr uint32_t hi = ((8u * SW) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | (uint64_t)(saddr >> 4);
}

template <int C>
struct Geo {
    static constexpr int SW = 2 * C;
    static constexpr int N3 = 3 * C;
    static constexpr int W_ROW = (N3 * SW + 1023) & ~1023;       // one kernel row of weights [3C x C]
    static constexpr int WS_ROW = (N3 * 32 + 1023) & ~1023;      // stem kernel row / bias tile [3C x 16]
    static constexpr int W_BUF = 3 * W_ROW + WS_ROW;             // a layer: 3 kernel rows + bias tile
};

template <int C>
__global__ void __launch_bounds__(THREADS, 1) k_trunk_p16(const __grid_constant__ CUtensorMap tmW_stem,
                                                          const __grid_constant__ CUtensorMap tmW_conv,
                                                          const __grid_constant__ CUtensorMap tmBias, Params P)
{
    using GE = Geo<C>;
    constexpr int SW = GE::SW, N3 = GE::N3, W_ROW = GE::W_ROW, WS_ROW = GE::WS_ROW, W_BUF = GE::W_BUF;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int rows = LEAD + 128 * P.T + 16;
    const int buf_bytes = (rows * SW + 1023) & ~1023;
    uint8_t *X = smem, *Y = smem + buf_bytes;
    uint8_t *W = smem + 2 * buf_bytes;                          // two layers' weights: layer n lives in buffer n & 1
    uint8_t *Id = W + 2 * W_BUF;                                // identity into the dx = 0 block [3C x C]
    uint8_t *Ones = Id + W_ROW;                                 // 128 rows x 16 ones
    uint64_t *bars = reinterpret_cast<uint64_t *>(Ones + 4096);
    uint64_t *tdone = bars;                                     // [t]
    uint64_t *aready = bars + MAX_TILES;                        // [layer & 1][t]
    uint64_t *wfull = bars + 3 * MAX_TILES, *wfree = bars + 3 * MAX_TILES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * MAX_TILES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + P.n_conv;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < P.T * N3) tmem_cols <<= 1;

    // ---- prologue independent of the previous kernel's output ----
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW_stem);
        tma_prefetch_desc(&tmW_conv);
        tma_prefetch_desc(&tmBias);
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    for (int i = threadIdx.x; i < 2 * buf_bytes / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < W_ROW / 16; i += THREADS) reinterpret_cast<uint4 *>(Id)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 4096 / 16; i += THREADS)
        reinterpret_cast<uint4 *>(Ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);

    griddep_wait();
    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    // positions per group from the live count: fewest 128-row tiles on the busiest CTA, ties to the larger group
    int G = 1;
    {
        int best = 0x7FFFFFFF;
        for (int g = 1; g <= P.G; ++g) {
            const int tiles = (POS_ROWS * g + 127) / 128;
            const int groups = (count + g - 1) / g;
            const int units = ((groups + (int)gridDim.x - 1) / (int)gridDim.x) * tiles;
            if (units <= best) { best = units; G = g; }
        }
    }
    const int T = (POS_ROWS * G + 127) / 128;
    const int n_groups = (count + G - 1) / G;
    const int issuers = min(ISSUERS, T);
    if (threadIdx.x == 0) {
        for (int t = 0; t < MAX_TILES; ++t) mbar_init(tdone + t, 1);
        for (int t = 0; t < 2 * MAX_TILES; ++t) mbar_init(aready + t, 4);
        for (int b = 0; b < 2; ++b) { mbar_init(wfull + b, 1); mbar_init(wfree + b, (uint32_t)issuers); }
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < C)            // rows C..2C-1 (the dx = 0 block) carry the identity
        *reinterpret_cast<__nv_bfloat16 *>(Id + swz<SW>((uint32_t)(C + threadIdx.x) * SW + (uint32_t)threadIdx.x * 2u)) =
            __float2bfloat16_rn(1.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);

    uint32_t lcount = 0;
    if (warp == 0 && lane == 0 && (int)blockIdx.x < n_groups) {
        mbar_expect_tx(wfull, 4u * N3 * 32u);
        for (int dy = 0; dy < 3; ++dy) tma_load_2d(W + dy * WS_ROW, &tmW_stem, wfull, 0, dy * N3);
        tma_load_2d(W + 3 * W_ROW, &tmBias, wfull, 0, 0);
    }

    const uint32_t x_addr = smem_u32(X), y_addr = smem_u32(Y);
    const uint64_t ones_desc = desc_of<32>(smem_u32(Ones));
    const uint32_t id_addr = smem_u32(Id);
    const uint32_t idesc = make_instr_desc(N3);
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);

    for (int group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int pos0 = group * G;
        // ---- observation planes -> channels 0..15 of the group's rows in Y ----
        if (warp > ISSUERS) {
            for (int idx = threadIdx.x - 32 * (1 + ISSUERS); idx < G * 81; idx += 32 * 4 * EPI_GROUPS) {
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
                uint4 o0, o1;
                o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                const uint32_t off = (uint32_t)(LEAD + POS_ROWS * g + 16 * (a / 9) + 1 + a % 9) * (uint32_t)SW;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off)) = o0;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off + 16)) = o1;
            }
            fence_async_smem();
        }
        __syncthreads();

        for (int l = 0; l < n_layers; ++l, ++lcount) {
            const uint32_t ph = lcount & 1u;
            const uint32_t wb = lcount & 1u, wph = (lcount >> 1) & 1u;
            const bool to_y = (l & 1) != 0;
            const uint32_t s_addr = to_y ? x_addr : y_addr;
            const uint32_t d_addr = to_y ? y_addr : x_addr;
            const bool last = l == n_layers - 1;
            const bool residual = l > 0 && !to_y;
            const bool stamp = P.timeline && blockIdx.x == 0 && group == 0 && lane == 0;
            long long *tl = P.timeline + 8 * l;
            if (warp == 0) {
                if (lane == 0) {
                    const bool more_layers = l + 1 < n_layers;
                    const bool more_groups = group + (int)gridDim.x < n_groups;
                    if (more_layers || more_groups) {
                        const uint32_t nb = wb ^ 1u;
                        uint8_t *Wn = W + nb * W_BUF;
                        if (lcount >= 1) mbar_wait(wfree + nb, ((lcount - 1) >> 1) & 1u);
                        if (more_layers) {
                            mbar_expect_tx(wfull + nb, 3u * N3 * (uint32_t)SW + N3 * 32u);
                            for (int dy = 0; dy < 3; ++dy)
                                tma_load_2d(Wn + dy * W_ROW, &tmW_conv, wfull + nb, 0, (l * 3 + dy) * N3);
                            tma_load_2d(Wn + 3 * W_ROW, &tmBias, wfull + nb, 0, (l + 1) * N3);
                        } else {
                            mbar_expect_tx(wfull + nb, 4u * N3 * 32u);
                            for (int dy = 0; dy < 3; ++dy) tma_load_2d(Wn + dy * WS_ROW, &tmW_stem, wfull + nb, 0, dy * N3);
                            tma_load_2d(Wn + 3 * W_ROW, &tmBias, wfull + nb, 0, 0);
                        }
                    }
                }
            } else if (warp <= ISSUERS) {
                const int e = warp - 1;
                if (e < issuers) {
                    if (stamp && warp == 1) tl[0] = clock64();
                    mbar_wait(wfull + wb, wph);
                    tc_fence_after();
                    if (stamp && warp == 1) tl[1] = clock64();
                    if (elect_one()) {
                        const uint32_t w_addr = smem_u32(W + wb * W_BUF);
                        const uint64_t bias_desc = desc_of<32>(w_addr + 3 * W_ROW);
                        for (int t = e; t < T; t += ISSUERS) {
                            if (l > 0) {
                                uint64_t *prev = aready + ((lcount - 1) & 1u) * MAX_TILES;
                                const uint32_t pph = ((lcount - 1) >> 1) & 1u;
                                if (t > 0) mbar_wait(prev + t - 1, pph);
                                mbar_wait(prev + t, pph);
                                if (t + 1 < T) mbar_wait(prev + t + 1, pph);
                                tc_fence_after();
                            }
                            const uint32_t d_tmem = tmem_base + (uint32_t)(t * N3);
                            const uint32_t a_tile = s_addr + (uint32_t)((LEAD + 128 * t) * SW);
                            umma_bf16(d_tmem, ones_desc, bias_desc, idesc, 0u);      // dx = 0 block starts as the bias
#pragma unroll 1
                            for (int dy = 0; dy < 3; ++dy) {
                                const int shift = (dy - 1) * 16 * SW;                // one slot up / down
                                if (l == 0) {
                                    umma_bf16(d_tmem, desc_of<SW>(a_tile + (uint32_t)shift),
                                              desc_of<32>(w_addr + (uint32_t)(dy * WS_ROW)), idesc, 1u);
                                } else {
#pragma unroll
                                    for (int k = 0; k < C / 16; ++k)
                                        umma_bf16(d_tmem, desc_of<SW>(a_tile + (uint32_t)shift + 32 * k),
                                                  desc_of<SW>(w_addr + (uint32_t)(dy * W_ROW) + 32 * k), idesc, 1u);
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
                    if (stamp && warp == 1) tl[2] = clock64();
                }
            } else {
                // ===== epilogue: out[m] = acc_-1[m-1] + acc_0[m] + acc_+1[m+1] (neighbours are in this warp), ReLU, store
                const int quarter = warp & 3, eg = (warp - 1 - ISSUERS) >> 2;
                for (int t = eg; t < T; t += EPI_GROUPS) {
                    const int q = 128 * t + 32 * quarter + lane;          // row inside the group
                    const int g = q / POS_ROWS, rem = q - POS_ROWS * g;
                    const int slot = rem >> 4, o = rem & 15;
                    const int row = pos0 + g;
                    const bool valid = g < G && row < count && slot < 9 && o >= 1 && o <= 9;
                    mbar_wait(tdone + t, ph);
                    tc_fence_after();
                    if (stamp && warp == 1 + ISSUERS && t == eg) tl[3] = clock64();
                    const uint32_t t_addr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(t * N3);
                    const uint32_t row_off = (uint32_t)(LEAD + q) * (uint32_t)SW;
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        uint32_t ra[32], rb[32], rc[32];
                        tmem_ld16(t_addr + (uint32_t)c0, ra);
                        tmem_ld16(t_addr + (uint32_t)(C + c0), rb);
                        tmem_ld16(t_addr + (uint32_t)(2 * C + c0), rc);
                        tmem_ld_wait();
                        float ov[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float up = __shfl_up_sync(0xFFFFFFFFu, __uint_as_float(ra[j]), 1);
                            const float dn = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(rc[j]), 1);
                            ov[j] = up + __uint_as_float(rb[j]) + dn;
                        }
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; j += 8) {
                                const uint32_t w0 = relu_pack(ov[j], ov[j + 1], zero2), w1 = relu_pack(ov[j + 2], ov[j + 3], zero2);
                                const uint32_t w2 = relu_pack(ov[j + 4], ov[j + 5], zero2), w3 = relu_pack(ov[j + 6], ov[j + 7], zero2);
                                if (last)
                                    *reinterpret_cast<uint4 *>(P.out + ((size_t)row * 81 + slot * 9 + (o - 1)) * C + c0 + j) =
                                        make_uint4(w0, w1, w2, w3);
                                else
                                    sts128(d_addr + swz<SW>(row_off + (uint32_t)(c0 + j) * 2u), w0, w1, w2, w3);
                            }
                        }
                    }
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(aready + wb * MAX_TILES + t);
                    if (stamp && warp == 1 + ISSUERS && t == eg) tl[4] = clock64();
                }
                if (stamp && warp == 1 + ISSUERS) tl[5] = clock64();
            }
        }
        if (P.timeline && blockIdx.x == 0 && group == 0 && threadIdx.x == 32) P.timeline[8 * n_layers] = clock64();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

template <int C>
static size_t smem_bytes(int T)
{
    using GE = Geo<C>;
    const size_t rows = LEAD + 128 * T + 16;
    const size_t buf = (rows * GE::SW + 1023) & ~(size_t)1023;
    return 1024 + 2 * buf + 2 * GE::W_BUF + GE::W_ROW + 4096 + (3 * MAX_TILES + 4) * 8 + 16;
}

}  // namespace p16

static size_t p16_smem(int C, int T) { return C == 16 ? p16::smem_bytes<16>(T) : p16::smem_bytes<32>(T); }

// positions / tiles per group: bounded by TMEM (T * 3C <= 512 columns) and shared memory
bool fused_p16_plan(int C, int &G, int &T)
{
    if (!(C == 16 || C == 32)) return false;
    const int t_tmem = std::min(p16::MAX_TILES, 512 / (3 * C));
    T = 0;
    for (int t = t_tmem; t >= 1; --t)
        if (p16_smem(C, t) <= 227 * 1024) { T = t; break; }
    if (T == 0) return false;
    G = (128 * T) / p16::POS_ROWS;
    return G >= 1;
}

size_t fused_p16_smem_bytes(int C, int T) { return p16_smem(C, T); }

int launch_trunk_p16(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G, int T,
                     int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                     __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st)
{
    p16::Params P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.G = G;
    P.T = T;
    P.n_conv = n_conv;
    P.in_planes = in_planes;
    P.states = states;
    P.obs = obs;
    P.out = out_flat;
    static long long *timeline_dev = nullptr;
    if (getenv("TACULT_FUSED_TIMELINE")) {
        if (!timeline_dev) cudaMalloc(&timeline_dev, 8 * 64 * sizeof(long long));
        cudaMemset(timeline_dev, 0, 8 * 64 * sizeof(long long));
        P.timeline = timeline_dev;
    }
    const size_t smem = p16_smem(C, T);
    const int grid = std::max(1, std::min(batch, sm_count));
    if (C == 16) {
        TC_CUDA(cudaFuncSetAttribute(p16::k_trunk_p16<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_CUDA(launch_chain(p16::k_trunk_p16<16>, grid, p16::THREADS, smem, st, w_stem, w_conv, bias_tiles, P));
    } else {
        TC_CUDA(cudaFuncSetAttribute(p16::k_trunk_p16<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_CUDA(launch_chain(p16::k_trunk_p16<32>, grid, p16::THREADS, smem, st, w_stem, w_conv, bias_tiles, P));
    }
    if (P.timeline) {
        long long h[8 * 64];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, timeline_dev, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[p16 timeline] layer: issuer-0 start | weights ready | issued | eg0 first tile done | its epilogue done | eg0 all done (cycles)\n");
        for (int l = 0; l <= n_conv; ++l)
            fprintf(stderr, "[p16 timeline] %2d: %7lld %7lld %7lld %7lld %7lld %7lld\n", l, h[8 * l] - h[0], h[8 * l + 1] - h[0],
                    h[8 * l + 2] - h[0], h[8 * l + 3] - h[0], h[8 * l + 4] - h[0], h[8 * l + 5] - h[0]);
        fprintf(stderr, "[p16 timeline] group end %7lld\n", h[8 * (n_conv + 1)] - h[0]);
    }
    return TC_OK;
}

}  // namespace tc
