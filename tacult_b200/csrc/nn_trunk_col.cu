// Column-tiled fused trunk for the 32-channel net: conv_in and every residual-block convolution of `_UtacNNet`
// (/root/reference/src/tacult/network.py:84-89, ResidualBlock 35-41) in one kernel, like k_trunk_fused (nn_fused.cu),
// but with HALF the shared-memory operand traffic per position and layer - the quantity that bounds a 32-channel
// implicit GEMM on this machine (a UMMA costs (4096 + 32 N) / 128 cycles of operand fetch).
//
// Layout.  A pass holds up to 12 positions.  UMMA tile j (j = 0..8) is BOARD COLUMN j of all of them: lane 10 g + r of
// the tile is cell (row r, column j) of position g, lane 10 g + 9 a zero pad row shared by positions g and g + 1, lanes
// 120..127 spare (zero).  With that layout
//   * a vertical tap (dy) is the same shared-memory tile read through a descriptor moved by dy rows, as before;
//   * a horizontal tap (dx) is ANOTHER TILE, SAME LANE.  So the three dx taps are stacked in N: the UMMA of input tile
//     i and kernel row dy multiplies the activations by the [3C x C] block (W[dy][dx=+1] | W[dy][0] | W[dy][dx=-1])
//     and accumulates into the 96 TMEM columns that hold output tiles i-1, i, i+1 side by side.  The sums over dx
//     happen in the accumulators; no lane ever needs another lane's value, and the epilogue is the old one
//     (one tcgen05.ld, ReLU, pack, store).
// Per tile and layer: 6 UMMAs of N = 96 (56 cycles each; N = 64 for the two edge columns, whose outer block has no
// output column) instead of 18 of N = 32 (40 each); 31 KB of operands per position and layer instead of 78 KB.
// Accumulators: 11 blocks of 32 TMEM columns (out tiles -1 and 9 are dummies).
//
// One thread issues every UMMA, in a fixed order: accumulating UMMAs of one thread into the same columns run back to
// back (56 cycles per N = 96 UMMA, measured: tools/probes/umma_issue_probe.cu) and the summation order is the same in
// every run, whereas two threads accumulating into overlapping columns serialise at ~250 cycles.  The issue code is
// straight-line per layer kind and weight buffer, every descriptor a (shared-memory base + compile-time constant):
// descriptors that travel through vector registers (R2UR) or a rolled tile loop cost 70-100 cycles per UMMA.  The whole
// walk over passes and layers sits inside ONE elected region (re-electing per layer cost ~600 cycles per layer).
//
// Bias and residual stay OFF the tensor pipe (as ones x bias-tile and identity-tile UMMAs they cost 9 x 46 and 18 x 58
// cycles per layer, a quarter of the kernel): the epilogue that drains an accumulator block stores the bias of the
// block's NEXT use back into it with tcgen05.st (TMEM writes run at 256 B/clk), so every conv UMMA simply accumulates;
// the residual is added by the epilogue from the shared-memory row it is about to overwrite.
//
// The passes of a CTA and their layers form one stream without a block-wide barrier: the stem of pass p+1 follows the
// last layer of pass p like any layer follows another (per-tile mbarriers), its observation planes having been encoded
// into a buffer of their own by two otherwise idle warps while pass p ran.  The planes come from ONE warp-cooperative
// legal-mask evaluation per position (rules_warp.cuh); per-cell scalar rules took a fifth of the kernel.
//
// What bounds it: shared-memory bandwidth.  Per tile and layer the tensor pipe fetches 42 KB of operands and the
// epilogue moves 8-16 KB (activation rows out, residual rows in) through the same 128 B/clk: 390-450 cycles, and the
// measured layer period is 3.85 k cycles for 9 tiles.  Roles: warp 0 streams the next layer's weights (TMA, double
// buffered), warp 1 issues, warps 2-3 encode planes, warps 4-15 are three epilogue groups of four (one warp per TMEM
// lane quarter; group e owns tiles j = e mod 3).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nn_bf16.cuh"
#include "nn_fused.cuh"
#include "rules.cuh"
#include "rules_warp.cuh"
#include "tcgen05.cuh"

#ifndef TC_COL_TILE_UNROLL
#define TC_COL_TILE_UNROLL 9
#endif

namespace tc {

namespace col {

constexpr int C = 32;
constexpr int SW = 2 * C;                // activation row = one swizzle span
constexpr int N3 = 3 * C;
constexpr int TILES = 9;                 // board columns
constexpr int PITCH = 10;                // lanes per position: 9 board rows + 1 zero row
constexpr int GMAX = 128 / PITCH;        // 12 positions per pass
constexpr int FIRST_EPI = 4;             // warp 0: TMA, warp 1: UMMA issue, warps 2-3: planes of the NEXT pass
constexpr int EPI_GROUPS = 3;            // epilogue groups of four warps (one per TMEM lane quarter); tiles j = group (mod 3)
constexpr int THREADS = 32 * (FIRST_EPI + 4 * EPI_GROUPS);
constexpr int LEAD = 16, TAIL = 16;      // zero rows around the nine tiles (1 KB each: tiles stay 1024-aligned)
constexpr int BUF_BYTES = (LEAD + 128 * TILES + TAIL) * SW;
constexpr int W_DY = N3 * SW;            // one kernel row of a conv layer: [3C x C] bf16, K-major, 64-byte swizzle
constexpr int WS_DY = N3 * 32;           // one kernel row of the stem:     [3C x 16]
constexpr int W_BUF = 3 * W_DY;
constexpr int BAR_STRIDE = 16;
constexpr int TILE_UNROLL = TC_COL_TILE_UNROLL;   // tiles of a layer issued as straight-line code

struct Params {
    const int *count_dev;
    int batch;
    int n_conv;              // 2 * residual blocks
    int in_planes;
    const uint32_t *states;  // packed positions, or
    const float *obs;        // dense planes [row][in_planes][81]
    const float *bias;       // [1 + n_conv][C] float32: stem, then every conv (BatchNorm folded)
    __nv_bfloat16 *out;      // compact [row][board column][board row][C] (fc1 input; NOT the [row][cell][C] of the other trunks)
    long long *timeline;     // optional clock64 marks of CTA 0
    int pdl_trigger;
};

__device__ __forceinline__ uint32_t swz(uint32_t off)       // byte offset inside a 1024-aligned buffer -> swizzled
{
    return off ^ (((off >> 7) & (uint32_t)(SW / 16 - 1)) << 4);
}
__device__ __forceinline__ uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }      // rows of 32 bytes

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

// Offsets of the shared-memory carve-up from its 1024-aligned base: compile-time constants, so that every UMMA
// descriptor of the issue loop is (uniform base + constant) and is built in the uniform datapath.  A descriptor that
// travels through a vector register costs an R2UR, and those were measured at ~35 cycles each in the issuing thread:
// 7 per tile made a tile cost 620 cycles instead of the 340-480 of its UMMAs (tools/probes/umma_issue_probe.cu).
constexpr uint32_t X_OFF = 0, Y_OFF = BUF_BYTES, W_OFF = 2 * BUF_BYTES;
constexpr uint32_t P_OFF = W_OFF + 2 * W_BUF;                    // observation planes of one pass: rows of 32 bytes (K = 16)
constexpr uint32_t PBUF_BYTES = (LEAD + 128 * TILES + TAIL) * 32;
constexpr uint32_t BARS_OFF = P_OFF + PBUF_BYTES;

// low word of a descriptor: (address >> 4); the base is 1024-aligned, so the shift distributes
template <int SWB>
__device__ __forceinline__ uint64_t desc_at(uint32_t base4, uint32_t off)
{
    constexpr uint32_t layout = SWB == 128 ? 2u : (SWB == 64 ? 4u : 6u);
    constexpr uint32_t hi = ((8u * SWB) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | (uint64_t)(base4 + (off >> 4));
}

// Every UMMA of one layer, issued by one thread in a fixed order.  KIND 0: stem (planes in Y -> X), 1: first conv of a
// block (X -> Y), 2: second conv (Y -> X).  WB: weight buffer.  Input tile i accumulates into the blocks of output tiles
// i-1, i, i+1 (which hold the layer's bias when it starts); output tile j is complete (tdone[j]) once input tile j+1 has
// been issued.
template <int KIND, int WB>
__device__ __forceinline__ void issue_layer(uint32_t base4, uint32_t tmem_base, uint64_t *tdone, uint64_t *prev, uint32_t pph,
                                            uint64_t *wfree_bar, uint64_t *next_wfull, uint32_t next_wph, bool first_of_kernel,
                                            uint64_t *pfree)
{
    constexpr uint32_t SRC = KIND == 0 ? P_OFF : (KIND == 1 ? X_OFF : Y_OFF);
    constexpr uint32_t AROW = KIND == 0 ? 32u : (uint32_t)SW;       // bytes per row of the A operand
    constexpr uint32_t WL = W_OFF + WB * W_BUF;
    constexpr uint32_t BROW = KIND == 0 ? 32u : (uint32_t)SW;       // bytes per row of the weight tile
    const uint32_t idesc96 = make_instr_desc(N3), idesc64 = make_instr_desc(2 * C);
#pragma unroll(TILE_UNROLL)
    for (int i = 0; i < TILES; ++i) {
        if (KIND != 0 || !first_of_kernel) {
            // the previous layer's output tile i is the A operand (not for a stem: its planes were waited for by the
            // caller); its blocks i (waited for by tile i-1) and i+1 must have been drained and re-initialised.  The stem
            // of a later pass follows the last layer of the pass before it like any other layer.
            if (i == 0) mbar_wait(prev, pph);
            if (i + 1 < TILES) mbar_wait(prev + i + 1, pph);
            tc_fence_after();
        }
        // the next layer's weights landed long ago: take their wait off the layer boundary
        if (i == TILES - 2 && next_wfull) mbar_wait(next_wfull, next_wph);
        const uint32_t tile = (uint32_t)LEAD * AROW + 128u * AROW * (uint32_t)i;
        // edge tiles: the outer block would go to a column that does not exist (N = 64, two blocks)
        const uint32_t d = tmem_base + (uint32_t)(C * i) + (i == 0 ? (uint32_t)C : 0u);
        const uint32_t wskip = i == 0 ? (uint32_t)C * BROW : 0u;
        const uint32_t idesc = (i == 0 || i == TILES - 1) ? idesc64 : idesc96;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const uint32_t a = SRC + tile + (uint32_t)dy * AROW - AROW;
            if (KIND == 0) {
                umma_bf16(d, desc_at<32>(base4, a), desc_at<32>(base4, WL + (uint32_t)(dy * W_DY) + wskip), idesc, 1u);
            } else {
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                    umma_bf16(d, desc_at<SW>(base4, a + 32u * k), desc_at<SW>(base4, WL + (uint32_t)(dy * W_DY) + wskip + 32u * k), idesc, 1u);
            }
        }
        if (i >= 1) umma_commit(tdone + i - 1);
    }
    umma_commit(tdone + TILES - 1);
    umma_commit(wfree_bar);
    if (KIND == 0) umma_commit(pfree);        // the planes buffer has been consumed: the next pass's planes may be written
}

__global__ void __launch_bounds__(THREADS, 1) k_trunk_col(const __grid_constant__ CUtensorMap tmW_stem,
                                                          const __grid_constant__ CUtensorMap tmW_conv, Params P)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *X = smem + X_OFF, *Y = smem + Y_OFF;
    uint8_t *W = smem + W_OFF;                                  // two layers' weights: layer n lives in buffer n & 1
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + BARS_OFF);
    uint64_t *tdone = bars;                                     // [j]: every UMMA into output tile j has completed
    uint64_t *aready = bars + BAR_STRIDE;                       // [layer & 1][j]: output tile j drained and written
    uint64_t *wfull = bars + 3 * BAR_STRIDE, *wfree = wfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wfree + 4);        // wfree + 2, + 3: `pready`, `pfree`

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + P.n_conv;

    // ---- prologue that depends on nothing the previous kernel wrote (programmatic dependent launch, common.cuh) ----
    griddep_launch(P.pdl_trigger == 1);
    long long *marks = (P.timeline && blockIdx.x == 0 && threadIdx.x == 0) ? P.timeline : nullptr;
    if (marks) marks[0] = clock64();
    long long *marks2 = (P.timeline && blockIdx.x == 0) ? P.timeline : nullptr;      // any thread of CTA 0
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW_stem);
        tma_prefetch_desc(&tmW_conv);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 32) {
        for (int j = 0; j < TILES; ++j) mbar_init(tdone + j, 1);
        for (int b = 0; b < 2; ++b)
            for (int j = 0; j < TILES; ++j) mbar_init(aready + b * BAR_STRIDE + j, 4);
        for (int b = 0; b < 2; ++b) { mbar_init(wfull + b, 1); mbar_init(wfree + b, 1); }
        mbar_init(wfree + 2, FIRST_EPI - 2);      // pready: one arrival per plane warp
        mbar_init(wfree + 3, 1);                  // pfree
        fence_barrier_init();
    }
    // The rows a shifted descriptor can reach from a live row without meaning to must read as zero for ever: the lead and
    // tail rows, the pad row after every position (lanes 9, 19, ...) and the spare lanes 120..127 of every tile.  Nothing
    // ever writes them.  Rows of positions a pass does not use keep whatever they held: a GEMM row only sees its own
    // position, and those rows are neither stored nor copied out.
    for (int i = threadIdx.x; i < 2 * (LEAD + TAIL) * 4; i += THREADS) {
        const int buf = i / ((LEAD + TAIL) * 4), rem = i - buf * (LEAD + TAIL) * 4;
        const int row = rem >> 2, ch = rem & 3;
        const int abs_row = row < LEAD ? row : LEAD + 128 * TILES + (row - LEAD);
        *reinterpret_cast<uint4 *>(smem + buf * BUF_BYTES + abs_row * SW + ch * 16) = make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < (int)PBUF_BYTES / 16; i += THREADS)          // the planes buffer: cleared whole (38 KB)
        reinterpret_cast<uint4 *>(smem + P_OFF)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * TILES * 20 * 4; i += THREADS) {
        const int ch = i & 3, k = (i >> 2) % 20, t = (i >> 2) / 20;               // t = buffer * TILES + tile
        const int buf = t / TILES, tile = t - buf * TILES;
        const int lane_row = k < GMAX ? PITCH * k + 9 : 120 + (k - GMAX);
        *reinterpret_cast<uint4 *>(smem + buf * BUF_BYTES + (LEAD + 128 * tile + lane_row) * SW + ch * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);
    // the stem's weights, and its bias as the initial value of the accumulator blocks (later passes and layers find the
    // bias the epilogue left there)
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(wfull, 3u * WS_DY);
        for (int dy = 0; dy < 3; ++dy) tma_load_2d(W + dy * W_DY, &tmW_stem, wfull, 0, dy * N3);
    }
    if (warp >= FIRST_EPI && warp < FIRST_EPI + 4) {
        uint32_t b0[32];
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(P.bias + c));
            b0[c] = __float_as_uint(v.x); b0[c + 1] = __float_as_uint(v.y); b0[c + 2] = __float_as_uint(v.z); b0[c + 3] = __float_as_uint(v.w);
        }
        for (int j = 0; j < TILES; ++j) tmem_st32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(C * (j + 1)), b0);
        tmem_st_wait();
        tc_fence_before();
    }

    // ---- from here on the leaf count and the leaf positions written by the descent kernel are read ----
    if (marks) marks[1] = clock64();
    griddep_wait();
    if (marks) marks[2] = clock64();
    // The live positions are dealt out round-robin: CTA b owns rows b, b + grid, b + 2 grid, ...  Which rows those are does
    // not depend on the leaf count, so the first pass's positions are fetched together WITH the count instead of after it
    // (one L2 round trip less on the critical path); rows beyond the count are dropped when it arrives.
    const int grid = (int)gridDim.x;
    uint32_t first_word = 0u;
    if (P.states && warp < GMAX && lane < 8 && (int)blockIdx.x + warp * grid < P.batch)
        first_word = P.states[8 * (size_t)((int)blockIdx.x + warp * grid) + lane];
    int count = P.batch;
    if (P.count_dev) count = min(count, *P.count_dev);
    const int my_n = count > (int)blockIdx.x ? (count - (int)blockIdx.x - 1) / grid + 1 : 0;
    const int n_pass = (my_n + GMAX - 1) / GMAX;
    auto row_of = [&](int k) { return (int)blockIdx.x + k * grid; };          // k-th position of this CTA
    if (n_pass == 0 && threadIdx.x == 0) {
        griddep_launch(P.pdl_trigger == 2);
        mbar_wait(wfull, 0u);                   // nothing to do: let the stem weights land before the CTA retires
    }
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    uint8_t *Pl = smem + P_OFF;
    uint64_t *pready = wfree + 2;      // planes of pass p >= 1 are in the planes buffer (completes once per such pass)
    uint64_t *pfree = wfree + 3;       // the stem of a pass has been consumed by the tensor pipe (completes once per pass)

    // Planes of one position -> channels 0..15 of its 81 rows of the planes buffer, by one warp: the legal mask is
    // computed ONCE, by the warp (rules_warp.cuh), then lane writes cells lane, lane + 32, lane + 64 (per-cell scalar
    // rules made this phase 8.7 k cycles per pass, a fifth of the kernel).
    auto encode_position = [&](int g, uint32_t w) {          // w: lane k < 8 holds word k of the packed position
        uint32_t wd[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) wd[k] = __shfl_sync(0xFFFFFFFFu, w, k);
        const Pos p = unpack(wd);
        uint32_t am[3], mm[3], oo[3];
        int status;
        warp_legal_mask<true>(p, lane, am, status, mm, oo);
        // lane -> cell in COLUMN-major order (consecutive lanes = consecutive board rows of one tile, 32 bytes apart): in
        // action order consecutive lanes are consecutive tiles, 4 KB apart on the same banks - 26 wavefronts per store
        // instead of 4-8 (ncu source page), 2.5 k wavefront-cycles per SM and launch stolen from the operand fetch
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int idx = lane + 32 * q;
            if (idx < 81) {
                const int cc = (idx * 57) >> 9, r = idx - 9 * cc;       // idx / 9, idx % 9
                const int a = 9 * r + cc, wsel = a >> 5, bit = a & 31;
                const uint32_t wm = wsel == 0 ? mm[0] : (wsel == 1 ? mm[1] : mm[2]);
                const uint32_t wo = wsel == 0 ? oo[0] : (wsel == 1 ? oo[1] : oo[2]);
                const uint32_t wl = wsel == 0 ? am[0] : (wsel == 1 ? am[1] : am[2]);
                const bool mine = (wm >> bit) & 1u, occ = (wo >> bit) & 1u, legal = (wl >> bit) & 1u;
                const uint32_t off = (uint32_t)(LEAD + 128 * cc + PITCH * g + r) * 32u;
                // bf16 1.0 = 0x3F80: planes mine, theirs | legal, ones; channels 4..15 are zero
                *reinterpret_cast<uint4 *>(Pl + swz32(off)) =
                    make_uint4((mine ? 0x3F80u : 0u) | ((occ && !mine) ? 0x3F800000u : 0u), (legal ? 0x3F80u : 0u) | 0x3F800000u, 0u, 0u);
                *reinterpret_cast<uint4 *>(Pl + swz32(off + 16)) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    // dense planes (tc_forward_obs_dev): consecutive threads write consecutive rows of one tile
    auto encode_dense = [&](int k0, int G, int tid, int nthreads) {       // k0: local index of the pass's first position
        for (int idx = tid; idx < G * 81; idx += nthreads) {
            const int cc = idx / (9 * G), gr = idx - 9 * G * cc;
            const int g = gr / 9, r = gr - 9 * g;
            const int a = 9 * r + cc;
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = c < P.in_planes ? P.obs[((size_t)row_of(k0 + g) * P.in_planes + c) * 81 + a] : 0.f;
            uint4 o0, o1;
            o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
            o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
            o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
            o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
            const uint32_t off = (uint32_t)(LEAD + 128 * cc + PITCH * g + r) * 32u;
            *reinterpret_cast<uint4 *>(Pl + swz32(off)) = o0;
            *reinterpret_cast<uint4 *>(Pl + swz32(off + 16)) = o1;
        }
    };

    // ---- planes of the first pass: every warp helps ----
    {
        const int G0 = min(GMAX, my_n);
        if (P.states) {
            if (warp < G0) encode_position(warp, first_word);
        } else {
            encode_dense(0, G0, (int)threadIdx.x, THREADS);
        }
        fence_async_smem();
        __syncthreads();
        if (marks) marks[3] = clock64();
    }

    // From here on there is no block-wide barrier until the kernel ends.  The passes and their layers form ONE stream:
    // input tile i of a layer needs the activations of output tile i of the layer before it and the accumulator blocks
    // of output tiles i-1, i, i+1 drained and re-initialised, nothing else - and the stem of pass p+1 follows the last
    // layer of pass p in exactly that way, its planes having been encoded into their own buffer by warps 2-3 while pass
    // p was running.  Every role walks the stream on its own (lc = layers executed so far: the mbarrier phases); the
    // issuer's whole walk sits inside ONE elected region - re-electing and re-converging the warp per layer cost ~600
    // cycles per layer.  Layer 0: planes -> X;  odd l (first conv of a block): X -> Y;  even l > 0: Y -> X, + residual X.
    if (warp == 0) {
        if (lane == 0) {
            for (int pass = 0; pass < n_pass; ++pass)
                for (int l = 0; l < n_layers; ++l) {
                    const uint32_t lc = (uint32_t)(pass * n_layers + l);
                    if (l == n_layers - 1 && pass == n_pass - 1) griddep_launch(P.pdl_trigger == 2);
                    const bool more_layers = l + 1 < n_layers;
                    const bool more_passes = pass + 1 < n_pass;
                    if (more_layers || more_passes) {
                        // stream the NEXT layer's weights into the other buffer once the layer before this one (its
                        // previous user) has been consumed by the tensor pipe
                        const uint32_t nb = (lc & 1u) ^ 1u;
                        uint8_t *Wn = W + nb * W_BUF;
                        if (lc >= 1) mbar_wait(wfree + nb, ((lc - 1) >> 1) & 1u);
                        if (more_layers) {
                            mbar_expect_tx(wfull + nb, 3u * W_DY);
                            for (int dy = 0; dy < 3; ++dy) tma_load_2d(Wn + dy * W_DY, &tmW_conv, wfull + nb, 0, (l * 3 + dy) * N3);
                        } else {
                            mbar_expect_tx(wfull + nb, 3u * WS_DY);
                            for (int dy = 0; dy < 3; ++dy) tma_load_2d(Wn + dy * W_DY, &tmW_stem, wfull + nb, 0, dy * N3);
                        }
                    }
                }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t base4 = smem_u32(smem) >> 4;
            for (int pass = 0; pass < n_pass; ++pass) {
                if (marks2 && pass < 4) marks2[150 + pass] = clock64();
                for (int l = 0; l < n_layers; ++l) {
                    const uint32_t lc = (uint32_t)(pass * n_layers + l);
                    const uint32_t wb = lc & 1u;
                    if (lc == 0) mbar_wait(wfull + wb, 0u);              // later layers: waited for inside the layer before
                    if (l == 0 && pass > 0) mbar_wait(pready, (uint32_t)(pass - 1) & 1u);
                    tc_fence_after();
                    if (marks2 && lc < 18) marks2[16 + 6 * lc] = clock64();
                    uint64_t *prev = aready + ((lc - 1) & 1u) * BAR_STRIDE;
                    const uint32_t pph = ((lc - 1) >> 1) & 1u;
                    uint64_t *nwf = (l + 1 < n_layers || pass + 1 < n_pass) ? wfull + (wb ^ 1u) : nullptr;
                    const uint32_t nwph = ((lc + 1) >> 1) & 1u;
                    if (l == 0) {
                        if (wb) issue_layer<0, 1>(base4, tmem_base, tdone, prev, pph, wfree + 1, nwf, nwph, lc == 0, pfree);
                        else issue_layer<0, 0>(base4, tmem_base, tdone, prev, pph, wfree, nwf, nwph, lc == 0, pfree);
                    } else if (l & 1) {
                        if (wb) issue_layer<1, 1>(base4, tmem_base, tdone, prev, pph, wfree + 1, nwf, nwph, false, pfree);
                        else issue_layer<1, 0>(base4, tmem_base, tdone, prev, pph, wfree, nwf, nwph, false, pfree);
                    } else {
                        if (wb) issue_layer<2, 1>(base4, tmem_base, tdone, prev, pph, wfree + 1, nwf, nwph, false, pfree);
                        else issue_layer<2, 0>(base4, tmem_base, tdone, prev, pph, wfree, nwf, nwph, false, pfree);
                    }
                    if (marks2 && lc < 18) marks2[17 + 6 * lc] = clock64();
                }
            }
        }
    } else if (warp < FIRST_EPI) {
        // ===== planes of pass p >= 1, while pass p-1 runs: once the stem of pass p-1 has been consumed by the tensor pipe
        //       the planes buffer is free (a barrier of its own: a parity wait cannot skip phases of a shared one) =====
        for (int pass = 1; pass < n_pass; ++pass) {
            mbar_wait(pfree, (uint32_t)(pass - 1) & 1u);
            const int k0 = pass * GMAX;
            const int G = min(GMAX, my_n - k0);
            if (P.states) {
                for (int g = warp - 2; g < G; g += FIRST_EPI - 2)
                    encode_position(g, lane < 8 ? P.states[8 * (size_t)row_of(k0 + g) + lane] : 0u);
            } else {
                encode_dense(k0, G, (int)threadIdx.x - 64, 32 * (FIRST_EPI - 2));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(pready);
        }
    } else {
        // ===== epilogue: (+ residual) ReLU, bf16 pack, store; the drained block gets the bias of its next use =====
        const int quarter = warp & 3, egroup = (warp - FIRST_EPI) >> 2;
        const int i = 32 * quarter + lane;                       // lane of every tile = (position, board row)
        const int g = i / PITCH, r = i - PITCH * g;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int k0 = pass * GMAX;
            const int G = min(GMAX, my_n - k0);
            const bool valid = g < G && r < 9;
            const size_t out_row = (size_t)row_of(k0 + g);
            for (int l = 0; l < n_layers; ++l) {
                const uint32_t lc = (uint32_t)(pass * n_layers + l);
                uint8_t *dst = (l & 1) ? Y : X;
                const bool residual = l > 0 && !(l & 1);             // second conv of a block: + the block's input, in X
                const bool last = l == n_layers - 1;
                uint32_t nb[32];                                     // bias of the next layer (the next pass's stem after the last)
                {
                    const float *bsrc = P.bias + (l + 1 < n_layers ? (l + 1) * C : 0);
#pragma unroll
                    for (int c = 0; c < C; c += 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(bsrc + c));
                        nb[c] = __float_as_uint(v.x); nb[c + 1] = __float_as_uint(v.y);
                        nb[c + 2] = __float_as_uint(v.z); nb[c + 3] = __float_as_uint(v.w);
                    }
                }
                for (int j = egroup; j < TILES; j += EPI_GROUPS) {
                    const uint32_t t_addr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(C * (j + 1));
                    mbar_wait(tdone + j, lc & 1u);
                    tc_fence_after();
                    uint32_t acc[32];
                    tmem_ld32(t_addr, acc);
                    tmem_ld_wait();
                    tmem_st32(t_addr, nb);
                    if (valid) {
                        const uint32_t row_off = (uint32_t)(LEAD + 128 * j + i) * (uint32_t)SW;
#pragma unroll
                        for (int c = 0; c < C; c += 8) {
                            uint4 *cell = reinterpret_cast<uint4 *>(dst + swz(row_off + (uint32_t)c * 2u));
                            if (residual) {
                                const uint4 x = *reinterpret_cast<const uint4 *>(X + swz(row_off + (uint32_t)c * 2u));
                                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    acc[c + 2 * q] = __float_as_uint(__uint_as_float(acc[c + 2 * q]) + __uint_as_float(xs[q] << 16));
                                    acc[c + 2 * q + 1] = __float_as_uint(__uint_as_float(acc[c + 2 * q + 1]) + __uint_as_float(xs[q] & 0xFFFF0000u));
                                }
                            }
                            uint4 ov;
                            ov.x = relu_pack(acc[c], acc[c + 1], zero2);
                            ov.y = relu_pack(acc[c + 2], acc[c + 3], zero2);
                            ov.z = relu_pack(acc[c + 4], acc[c + 5], zero2);
                            ov.w = relu_pack(acc[c + 6], acc[c + 7], zero2);
                            // The fc1 input is column-major per position ([row][board column][board row][C], the K order
                            // fc1's weights are uploaded in for this kernel): the nine rows of a position's column are 576
                            // contiguous bytes, written by nine adjacent lanes.  (The 62 KB a pass writes leave every SM in
                            // the same 2 k cycles; sending them through shared memory and the TMA engine instead - 576-byte
                            // bulk copies - was measured slower, 44.4 vs 45.1 M sims/s.)
                            if (last) *reinterpret_cast<uint4 *>(P.out + (out_row * 81 + 9 * j + r) * C + c) = ov;
                            else *cell = ov;
                        }
                    }
                    // rows written through the generic proxy are read by the next layer's UMMAs (async proxy); the
                    // accumulator block holds its next bias
                    tmem_st_wait();
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(aready + (lc & 1u) * BAR_STRIDE + j);
                    if (marks2 && lc < 18 && warp == FIRST_EPI && lane == 0) marks2[(j == 0 ? 18 : 19) + 6 * lc] = clock64();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (marks) { marks[11] = clock64(); marks[12] = n_pass; marks[13] = my_n; }
}

static size_t smem_bytes() { return 1024 + 2 * (size_t)BUF_BYTES + 2 * (size_t)W_BUF + PBUF_BYTES + (3 * BAR_STRIDE + 6) * 8 + 16; }

}  // namespace col

bool trunk_col_supported(int C, int in_planes) { return C == col::C && in_planes <= 16; }

// Weights of one layer as three [3C x K] blocks, one per kernel row ky: rows 0..C-1 hold kx = 2 (dx = +1, accumulates
// into output tile i-1), rows C..2C-1 kx = 1, rows 2C..3C-1 kx = 0.  `w` is [co][ci][3][3] (PyTorch), K = ci padded to
// k_pad columns.
void trunk_col_pack(const float *w, int c_out, int c_in, int k_pad, __nv_bfloat16 *dst)
{
    for (int ky = 0; ky < 3; ++ky)
        for (int b = 0; b < 3; ++b)
            for (int co = 0; co < c_out; ++co)
                for (int ci = 0; ci < k_pad; ++ci)
                    dst[(((size_t)ky * 3 + b) * c_out + co) * k_pad + ci] =
                        __float2bfloat16_rn(ci < c_in ? w[((size_t)co * c_in + ci) * 9 + ky * 3 + (2 - b)] : 0.f);
}

int launch_trunk_col(const CUtensorMap &w_stem, const CUtensorMap &w_conv, const float *bias_all, int n_conv,
                     int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                     __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st)
{
    col::Params P{};
    P.count_dev = count_dev;
    P.batch = batch;
    P.n_conv = n_conv;
    P.in_planes = in_planes;
    P.states = states;
    P.obs = obs;
    P.out = out_flat;
    P.bias = bias_all;
    P.pdl_trigger = pdl_trigger_enabled();
    static long long *timeline_dev = nullptr;
    if (getenv("TACULT_FUSED_TIMELINE")) {
        if (!timeline_dev) { cudaMalloc(&timeline_dev, 160 * sizeof(long long)); cudaMemset(timeline_dev, 0, 160 * sizeof(long long)); }
        P.timeline = timeline_dev;
    }
    const size_t smem = col::smem_bytes();
    const int grid = std::max(1, std::min(batch, sm_count));
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        TC_CUDA(cudaFuncSetAttribute(col::k_trunk_col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    TC_CUDA(launch_chain(col::k_trunk_col, grid, col::THREADS, smem, st, w_stem, w_conv, P));
    if (P.timeline) {
        long long m[160];
        cudaStreamSynchronize(st);
        cudaMemcpy(m, timeline_dev, sizeof m, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[col marks] CTA 0, %lld positions in %lld passes; cycles since kernel entry: prologue done %lld | "
                "dependency wait over %lld | planes of pass 0 done %lld | issuer starts pass 0 at %lld, pass 1 at %lld, pass 2 at %lld | "
                "exit %lld\n",
                m[13], m[12], m[1] - m[0], m[2] - m[0], m[3] - m[0], m[150] - m[0], m[12] > 1 ? m[151] - m[0] : 0,
                m[12] > 2 ? m[152] - m[0] : 0, m[11] - m[0]);
        fprintf(stderr, "[col layers] layer of the stream: issue start | issue end | this group's epilogue of tile 0 | of its last tile "
                "(cycles since kernel entry)\n");
        for (int l = 0; l < (int)m[12] * (n_conv + 1) && l < 18; ++l)
            fprintf(stderr, "[col layers] %2d: %7lld %7lld %7lld %7lld\n", l, m[16 + 6 * l] - m[0], m[17 + 6 * l] - m[0],
                    m[18 + 6 * l] - m[0], m[19 + 6 * l] - m[0]);
    }
    return TC_OK;
}

}  // namespace tc
