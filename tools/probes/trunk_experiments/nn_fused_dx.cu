// Fused trunk, second formulation: each A fetch feeds the three horizontal taps.
//
// In k_trunk_fused (nn_fused.cu) the tensor pipe is bound by operand reads from shared memory: with N = C = 32 a
// UMMA moves 5 KB for 65 k MACs (measured ~41 cycles per UMMA against 16 of math).  Here the three taps of one
// kernel row (dx = -1, 0, +1) share one A tile: for each dy the UMMA multiplies the activation rows
// [m + 10 dy] by the [3C x C] weight block of that row, so the accumulator of a tile has 3C columns
//     acc_dx[m] = sum_dy X[m + 10 dy] * W[dy][dx]          and          out[m] = acc_-1[m-1] + acc_0[m] + acc_+1[m+1].
// The +-1-row combine happens in the epilogue: shuffles inside a warp, a 1 KB shared-memory exchange between the
// four warps of a tile; tiles overlap by two rows (126 useful rows per 128-lane tile) so no tile needs another one.
// Operand traffic per tile and layer drops from 18 x 5 KB to 6 x 7 KB.  Bias and residual still ride on the tensor
// pipe (ones x bias tile, X x identity into the dx = 0 block).  TMEM holds up to four 3C-column accumulators, one per
// (issuer warp, epilogue group) pair with full/empty mbarriers; 4 issuer warps, 16 epilogue warps, 1 TMA warp.
//
// STATUS (measured on B200, round 1): exact, but SLOWER than k_trunk_fused - opt-in with TACULT_FUSED_DX=1.  The
// tensor side does get ~2x cheaper (10 UMMAs of ~42 cycles per tile against 21), but the +-1-row combine costs 64
// warp shuffles per warp and tile and the epilogue becomes shuffle/MIO-bound (~600 cycles per tile against ~420 of
// UMMA time; 178 us against 104 us for 4096 positions).  tcgen05.shift (rows down by one inside each 32-lane block)
// was probed as the alternative: ~48 cycles per 8-column strip, 12 strips per tile - slower still
// (tools/probes/tmem_shift_probe.cu).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "tcgen05.cuh"

namespace tc {

namespace dx {

constexpr int ISSUERS = 4;
constexpr int EPI_GROUPS = 4;            // epilogue groups of 4 warps (one per TMEM lane quarter)
constexpr int THREADS = 32 * (1 + ISSUERS + 4 * EPI_GROUPS);
constexpr int LEAD = 16;
constexpr int TILE = 126;                // useful rows per 128-lane tile (lanes 1..126)
constexpr int MAX_SLOTS = 8;

struct Params {
    const int *count_dev;
    int batch;
    int G, T;                // maxima the shared-memory carve-up was sized for
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

__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)
{
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
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
    // one accumulator slot per (issuer warp, epilogue group) pair: both sides see every phase of their slot's
    // mbarriers, which the parity wait requires
    static constexpr int SLOTS = (512 / N3) < ISSUERS ? (512 / N3) : ISSUERS;
    static constexpr int W_ROW = (N3 * SW + 1023) & ~1023;       // weights of one kernel row  [3C x C]
    static constexpr int WS_ROW = (N3 * 32 + 1023) & ~1023;      // stem / bias tile           [3C x 16]
    static constexpr int XCH = MAX_SLOTS * 4 * 2 * C * 4;        // exchange: per slot, per warp, (a of lane 31, c of lane 0)
};

template <int C>
__global__ void __launch_bounds__(THREADS, 1) k_trunk_dx(const __grid_constant__ CUtensorMap tmW_stem,
                                                         const __grid_constant__ CUtensorMap tmW_conv,
                                                         const __grid_constant__ CUtensorMap tmBias, Params P)
{
    using GE = Geo<C>;
    constexpr int SW = GE::SW, N3 = GE::N3, SLOTS = GE::SLOTS, W_ROW = GE::W_ROW, WS_ROW = GE::WS_ROW;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int rows = LEAD + TILE * P.T + 16;
    const int buf_bytes = (rows * SW + 1023) & ~1023;
    uint8_t *X = smem, *Y = smem + buf_bytes;
    uint8_t *W = smem + 2 * buf_bytes;                          // 3 kernel rows of the current layer
    uint8_t *Bt = W + 3 * W_ROW;                                // bias tile of the current layer
    uint8_t *Id = Bt + WS_ROW;                                  // identity into the dx = 0 block  [3C x C]
    uint8_t *Ones = Id + W_ROW;                                 // 128 rows x 16 ones
    float *Xch = reinterpret_cast<float *>(Ones + 4096);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(Xch) + GE::XCH);
    uint64_t *tfull = bars, *tempty = bars + MAX_SLOTS, *wfull = bars + 2 * MAX_SLOTS, *wfree = bars + 2 * MAX_SLOTS + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_SLOTS + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    const int waves = max(1, (count + (int)gridDim.x * P.G - 1) / ((int)gridDim.x * P.G));
    const int G = max(1, min(P.G, (count + (int)gridDim.x * waves - 1) / ((int)gridDim.x * waves)));
    const int T = (100 * G + TILE - 1) / TILE;
    const int n_groups = (count + G - 1) / G;
    const int n_layers = 1 + P.n_conv;
    const int issuers = min(SLOTS, T);

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        mbar_init(wfull, 1);
        mbar_init(wfree, (uint32_t)issuers);
        fence_barrier_init();
        tma_prefetch_desc(&tmW_stem);
        tma_prefetch_desc(&tmW_conv);
        tma_prefetch_desc(&tmBias);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 2 * buf_bytes / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < W_ROW / 16; i += THREADS) reinterpret_cast<uint4 *>(Id)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 4096 / 16; i += THREADS)
        reinterpret_cast<uint4 *>(Ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    __syncthreads();
    if (threadIdx.x < C)            // rows C..2C-1 (the dx = 0 block) are the identity
        *reinterpret_cast<__nv_bfloat16 *>(Id + swz<SW>((uint32_t)(C + threadIdx.x) * SW + (uint32_t)threadIdx.x * 2u)) =
            __float2bfloat16_rn(1.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);

    uint32_t lcount = 0;       // layers executed so far by this CTA
    uint32_t use = 0;          // tiles this warp's accumulator slot has carried so far
    if (warp == 0 && lane == 0 && (int)blockIdx.x < n_groups) {
        mbar_expect_tx(wfull, 4u * N3 * 32u);
        for (int dy = 0; dy < 3; ++dy) tma_load_2d(W + dy * WS_ROW, &tmW_stem, wfull, 0, dy * N3);
        tma_load_2d(Bt, &tmBias, wfull, 0, 0);
    }

    const uint32_t x_addr = smem_u32(X), y_addr = smem_u32(Y), w_addr = smem_u32(W);
    const uint64_t ones_desc = desc_of<32>(smem_u32(Ones)), bias_desc = desc_of<32>(smem_u32(Bt));
    const uint32_t id_addr = smem_u32(Id);
    const uint32_t idesc = make_instr_desc(N3);
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);

    for (int group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int pos0 = group * G;
        // observation planes -> channels 0..15 of the group's rows in Y
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
                const uint32_t off = (uint32_t)(LEAD + 100 * g + (a / 9) * 10 + a % 9) * (uint32_t)SW;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off)) = o0;
                *reinterpret_cast<uint4 *>(Y + swz<SW>(off + 16)) = o1;
            }
            fence_async_smem();
        }
        __syncthreads();

        for (int l = 0; l < n_layers; ++l, ++lcount) {
            const uint32_t ph = lcount & 1u;
            const bool to_y = (l & 1) != 0;
            const uint32_t s_addr = to_y ? x_addr : y_addr;
            uint8_t *dst = to_y ? Y : X;
            const bool last = l == n_layers - 1;
            const bool residual = l > 0 && !to_y;
            const bool stamp = P.timeline && blockIdx.x == 0 && group == 0;
            long long *tl = P.timeline + 32 * l;
            if (stamp && threadIdx.x == 32) tl[0] = clock64();
            if (warp == 0) {
                if (lane == 0) {
                    mbar_wait(wfree, ph);
                    const bool more_layers = l + 1 < n_layers;
                    const bool more_groups = group + (int)gridDim.x < n_groups;
                    if (more_layers) {
                        mbar_expect_tx(wfull, 3u * N3 * (uint32_t)SW + N3 * 32u);
                        for (int dy = 0; dy < 3; ++dy)
                            tma_load_2d(W + dy * W_ROW, &tmW_conv, wfull, 0, (l * 3 + dy) * N3);
                        tma_load_2d(Bt, &tmBias, wfull, 0, (l + 1) * N3);
                    } else if (more_groups) {
                        mbar_expect_tx(wfull, 4u * N3 * 32u);
                        for (int dy = 0; dy < 3; ++dy) tma_load_2d(W + dy * WS_ROW, &tmW_stem, wfull, 0, dy * N3);
                        tma_load_2d(Bt, &tmBias, wfull, 0, 0);
                    }
                }
            } else if (warp <= ISSUERS) {
                const int e = warp - 1;
                if (e < issuers) {
                    mbar_wait(wfull, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t slot = (uint32_t)e;
                        uint32_t u = use;
                        if (stamp && e == 0) tl[1] = clock64();
                        for (int t = e; t < T; t += SLOTS, ++u) {
                            mbar_wait(tempty + slot, (u & 1u) ^ 1u);        // accumulator slot drained by the epilogue
                            tc_fence_after();
                            if (stamp && e == 0 && t < 3 * SLOTS) tl[2 + 2 * (t / SLOTS)] = clock64();
                            const uint32_t d_tmem = tmem_base + slot * (uint32_t)N3;
                            // lane i of the tile is padded row 126 t - 1 + i of the group
                            const uint32_t a_tile = s_addr + (uint32_t)((LEAD + TILE * t - 1) * SW);
                            umma_bf16(d_tmem, ones_desc, bias_desc, idesc, 0u);
#pragma unroll 1
                            for (int dy = 0; dy < 3; ++dy) {
                                const int shift = (dy - 1) * 10 * SW;
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
                                    umma_bf16(d_tmem, desc_of<SW>(x_addr + (uint32_t)((LEAD + TILE * t - 1) * SW) + 32 * k),
                                              desc_of<SW>(id_addr + 32 * k), idesc, 1u);
                            }
                            umma_commit(tfull + slot);
                            if (stamp && e == 0 && t < 3 * SLOTS) tl[3 + 2 * (t / SLOTS)] = clock64();
                        }
                        umma_commit(wfree);
                    }
                    use += (uint32_t)((T - e + SLOTS - 1) / SLOTS);         // every lane keeps the count (any may be elected)
                    __syncwarp();
                }
            } else {
                // ===== epilogue: combine the three dx blocks across neighbouring rows, ReLU, pack, store =====
                const int ew = warp - 1 - ISSUERS, quarter = warp & 3, eg = ew >> 2;
                const uint32_t slot = (uint32_t)eg;
                for (int t = eg; eg < SLOTS && t < T; t += SLOTS, ++use) {
                    const int i = 32 * quarter + lane;                    // TMEM lane of this thread
                    const int q = TILE * t - 1 + i;                       // padded row inside the group
                    const int g = q >= 0 ? q / 100 : 0, rem = q - 100 * g;
                    const int r = rem / 10, cc = rem - 10 * r;
                    const int row = pos0 + g;
                    const bool valid = i >= 1 && i <= TILE && q >= 0 && g < G && row < count && rem < 90 && cc != 9;
                    mbar_wait(tfull + slot, use & 1u);
                    tc_fence_after();
                    const bool es = stamp && ew == 0 && lane == 0 && t < 3 * SLOTS;
                    long long *te = tl + 8 + 4 * (t / SLOTS);
                    if (es) te[0] = clock64();
                    const uint32_t t_addr = tmem_base + ((uint32_t)(32 * quarter) << 16) + slot * (uint32_t)N3;
                    const uint32_t xch = smem_u32(Xch);
                    const uint32_t xa = xch + (uint32_t)(((slot * 4 + quarter) * 2 + 0) * C * 4);      // a of this warp's lane 31
                    const uint32_t xc = xch + (uint32_t)(((slot * 4 + quarter) * 2 + 1) * C * 4);      // c of this warp's lane 0
                    // (1) publish the two rows the neighbouring warps need
                    {
                        constexpr int CW = C < 32 ? 16 : 32;
                        uint32_t ra[32], rc[32];
#pragma unroll
                        for (int c0 = 0; c0 < C; c0 += CW) {
                            if (CW == 32) {
                                tmem_ld32(t_addr + (uint32_t)c0, ra);
                                tmem_ld32(t_addr + (uint32_t)(2 * C + c0), rc);
                            } else {
                                tmem_ld16(t_addr + (uint32_t)c0, ra);
                                tmem_ld16(t_addr + (uint32_t)(2 * C + c0), rc);
                            }
                            tmem_ld_wait();
                            if (lane == 31) {
#pragma unroll
                                for (int j = 0; j < CW; j += 4) sts128(xa + (uint32_t)(c0 + j) * 4u, ra[j], ra[j + 1], ra[j + 2], ra[j + 3]);
                            }
                            if (lane == 0) {
#pragma unroll
                                for (int j = 0; j < CW; j += 4) sts128(xc + (uint32_t)(c0 + j) * 4u, rc[j], rc[j + 1], rc[j + 2], rc[j + 3]);
                            }
                        }
                    }
                    if (es) te[1] = clock64();
                    named_bar_sync(1 + eg, 128);
                    if (es) te[2] = clock64();
                    const uint32_t na = xch + (uint32_t)(((slot * 4 + (quarter > 0 ? quarter - 1 : 0)) * 2 + 0) * C * 4);
                    const uint32_t nc = xch + (uint32_t)(((slot * 4 + (quarter < 3 ? quarter + 1 : 3)) * 2 + 1) * C * 4);
                    const uint32_t row_off = (uint32_t)(LEAD + q) * (uint32_t)SW;
                    const uint32_t dst_addr = to_y ? y_addr : x_addr;
                    const int src_up = (lane + 31) & 31, src_dn = (lane + 1) & 31;
                    // (2) out[m] = acc_-1[m-1] + acc_0[m] + acc_+1[m+1], 16 channels at a time.  Lane 31 swaps its own
                    //     acc_-1 row (needed by nobody in this warp) for the previous warp's, lane 0 its acc_+1 row for the
                    //     next warp's; a rotation shuffle then serves every lane.
#pragma unroll
                    for (int c0 = 0; c0 < C; c0 += 16) {
                        uint32_t ra[32], rb[32], rc[32];
                        tmem_ld16(t_addr + (uint32_t)c0, ra);
                        tmem_ld16(t_addr + (uint32_t)(C + c0), rb);
                        tmem_ld16(t_addr + (uint32_t)(2 * C + c0), rc);
                        tmem_ld_wait();
                        if (lane == 31) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) lds128(na + (uint32_t)(c0 + j) * 4u, ra[j], ra[j + 1], ra[j + 2], ra[j + 3]);
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) lds128(nc + (uint32_t)(c0 + j) * 4u, rc[j], rc[j + 1], rc[j + 2], rc[j + 3]);
                        }
                        if (c0 + 16 >= C) {               // accumulator and exchange rows fully read: hand the slot back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty + slot);
                        }
                        float o[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float up = __uint_as_float(__shfl_sync(0xFFFFFFFFu, ra[j], src_up));
                            const float dn = __uint_as_float(__shfl_sync(0xFFFFFFFFu, rc[j], src_dn));
                            o[j] = up + __uint_as_float(rb[j]) + dn;
                        }
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; j += 8) {
                                uint4 ov;
                                ov.x = relu_pack(o[j], o[j + 1], zero2);
                                ov.y = relu_pack(o[j + 2], o[j + 3], zero2);
                                ov.z = relu_pack(o[j + 4], o[j + 5], zero2);
                                ov.w = relu_pack(o[j + 6], o[j + 7], zero2);
                                if (last)
                                    *reinterpret_cast<uint4 *>(P.out + ((size_t)row * 81 + r * 9 + cc) * C + c0 + j) = ov;
                                else
                                    sts128(dst_addr + swz<SW>(row_off + (uint32_t)(c0 + j) * 2u), ov.x, ov.y, ov.z, ov.w);
                            }
                        }
                    }
                    if (es) te[3] = clock64();
                }
                fence_async_smem();
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (stamp && threadIdx.x == 32) tl[24] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int C>
static size_t smem_bytes(int T)
{
    using GE = Geo<C>;
    const size_t rows = LEAD + TILE * T + 16;
    const size_t buf = (rows * GE::SW + 1023) & ~(size_t)1023;
    return 1024 + 2 * buf + 3 * GE::W_ROW + GE::WS_ROW + GE::W_ROW + 4096 + GE::XCH + (2 * MAX_SLOTS + 2) * 8 + 16;
}

}  // namespace dx

static size_t dx_smem(int C, int T)
{
    return C == 16 ? dx::smem_bytes<16>(T) : (C == 32 ? dx::smem_bytes<32>(T) : dx::smem_bytes<64>(T));
}

bool fused_dx_plan(int C, int &G, int &T)
{
    if (!(C == 16 || C == 32 || C == 64)) return false;
    T = 0;
    for (int t = 24; t >= 1; --t)
        if (dx_smem(C, t) <= 227 * 1024) { T = t; break; }
    if (T == 0) return false;
    G = (dx::TILE * T) / 100;
    return G >= 1;
}

// weights of conv layer l as three kernel rows of [3C x C] (row = dx*C + cout), K-major
void fused_dx_pack_weights(const std::vector<std::vector<float>> &conv_w, int C, std::vector<__nv_bfloat16> &out)
{
    const size_t L = conv_w.size();
    out.assign(std::max<size_t>(1, L) * 3 * 3 * C * C, __float2bfloat16_rn(0.f));
    for (size_t l = 0; l < L; ++l)
        for (int dy = 0; dy < 3; ++dy)
            for (int dxk = 0; dxk < 3; ++dxk)
                for (int co = 0; co < C; ++co)
                    for (int ci = 0; ci < C; ++ci)
                        out[(((l * 3 + dy) * 3 + dxk) * C + co) * C + ci] =
                            __float2bfloat16_rn(conv_w[l][((size_t)co * C + ci) * 9 + dy * 3 + dxk]);
}

// stem: [3 dy][3C rows][16 cols]
void fused_dx_pack_stem(const std::vector<float> &stem_w, int C, int in_planes, std::vector<__nv_bfloat16> &out)
{
    out.assign((size_t)3 * 3 * C * 16, __float2bfloat16_rn(0.f));
    for (int dy = 0; dy < 3; ++dy)
        for (int dxk = 0; dxk < 3; ++dxk)
            for (int co = 0; co < C; ++co)
                for (int ci = 0; ci < in_planes; ++ci)
                    out[((size_t)(dy * 3 + dxk) * C + co) * 16 + ci] =
                        __float2bfloat16_rn(stem_w[((size_t)co * in_planes + ci) * 9 + dy * 3 + dxk]);
}

// bias tiles [layer][3C rows][16]: only the dx = 0 block carries the bias (bf16 head + remainder)
void fused_dx_bias_tiles(const std::vector<float> &bias_all, int C, std::vector<__nv_bfloat16> &out)
{
    const size_t layers = bias_all.size() / C;
    out.assign(layers * 3 * C * 16, __float2bfloat16_rn(0.f));
    for (size_t l = 0; l < layers; ++l)
        for (int co = 0; co < C; ++co) {
            const float b = bias_all[l * C + co];
            __nv_bfloat16 hi = __float2bfloat16_rn(b);
            __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            out[((l * 3 + 1) * C + co) * 16] = hi;
            out[((l * 3 + 1) * C + co) * 16 + 1] = lo;
        }
}

int launch_trunk_dx(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G, int T,
                    int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                    __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st)
{
    dx::Params P{};
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
        if (!timeline_dev) {
            cudaMalloc(&timeline_dev, 32 * 64 * sizeof(long long));
            cudaMemset(timeline_dev, 0, 32 * 64 * sizeof(long long));
        }
        P.timeline = timeline_dev;
    }
    const size_t smem = dx_smem(C, T);
    const int grid = std::max(1, std::min(batch, sm_count));
    if (C == 16) {
        TC_CUDA(cudaFuncSetAttribute(dx::k_trunk_dx<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dx::k_trunk_dx<16><<<grid, dx::THREADS, smem, st>>>(w_stem, w_conv, bias_tiles, P);
    } else if (C == 32) {
        TC_CUDA(cudaFuncSetAttribute(dx::k_trunk_dx<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dx::k_trunk_dx<32><<<grid, dx::THREADS, smem, st>>>(w_stem, w_conv, bias_tiles, P);
    } else {
        TC_CUDA(cudaFuncSetAttribute(dx::k_trunk_dx<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dx::k_trunk_dx<64><<<grid, dx::THREADS, smem, st>>>(w_stem, w_conv, bias_tiles, P);
    }
    if (P.timeline) {
        static long long h[32 * 64];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, timeline_dev, sizeof h, cudaMemcpyDeviceToHost);
        for (int l = 0; l <= n_conv; ++l) {
            const long long *t = h + 32 * l, t0 = t[0];
            fprintf(stderr, "[dx timeline] %2d: weights %5lld |", l, t[1] - t0);
            for (int k = 0; k < 3; ++k)
                if (t[2 + 2 * k]) fprintf(stderr, " issue%d %5lld..%5lld", k, t[2 + 2 * k] - t0, t[3 + 2 * k] - t0);
            fprintf(stderr, " |");
            for (int k = 0; k < 3; ++k)
                if (t[8 + 4 * k])
                    fprintf(stderr, " epi%d full %5lld pub %5lld bar %5lld done %5lld", k, t[8 + 4 * k] - t0, t[9 + 4 * k] - t0,
                            t[10 + 4 * k] - t0, t[11 + 4 * k] - t0);
            fprintf(stderr, " | layer %5lld\n", t[24] - t0);
        }
        cudaMemset(timeline_dev, 0, 32 * 64 * sizeof(long long));
    }
    return TC_OK;
}

}  // namespace tc
