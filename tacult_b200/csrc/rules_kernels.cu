// Rules-engine entry points of the C ABI: batched make-move, legal mask, terminal status,
// observation planes and seeded playouts, all as bit-parallel integer kernels.
// Reference call sites replaced: /root/reference/src/tacult/utac_game.py:34-107,247-250 and
// /root/reference/src/tacult/base/utils.py:7-10 (which reach the external C++ `utac`).
#include "common.cuh"
#include "rules.cuh"

namespace tc {

__global__ void k_rules_step(int n, const uint32_t *__restrict__ in, const int8_t *__restrict__ actions,
                             uint32_t *__restrict__ out, int8_t *__restrict__ status, int *__restrict__ illegal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    const uint4 *src = reinterpret_cast<const uint4 *>(in + 8 * (size_t)i);
    uint4 a0 = src[0], a1 = src[1];
    w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w; w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
    Pos p = unpack(w);
    Summary s = summarize(p);
    uint32_t legal[3];
    legal_words(p, s, legal);
    int a = actions[i];
    if (a < 0 || a > 80 || !action_bit(legal, a)) {
        atomicAdd(illegal, 1);
        pack(p, w);
        status[i] = (int8_t)status_of(s);
    } else {
        Pos q = play(p, a);
        pack(q, w);
        status[i] = (int8_t)status_of(summarize(q));
    }
    uint4 *dst = reinterpret_cast<uint4 *>(out + 8 * (size_t)i);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// one warp per position: lane l owns actions l, l+32, l+64
__global__ void k_rules_mask_obs(int n, const uint32_t *__restrict__ in, uint8_t *__restrict__ mask,
                                 float *__restrict__ obs, int8_t *__restrict__ status)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n) return;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = in[8 * (size_t)warp + k];
    Pos p = unpack(w);
    Summary s = summarize(p);
    uint32_t legal[3];
    legal_words(p, s, legal);
    if (status && lane == 0) status[warp] = (int8_t)status_of(s);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int a = lane + 32 * q;
        if (a < 81) {
            bool lg = action_bit(legal, a);
            if (mask) mask[81 * (size_t)warp + a] = lg ? 1 : 0;
            if (obs) {
                bool mine = action_bit(p.mine, a), occ = action_bit(p.occ, a);
                float *o = obs + 324 * (size_t)warp;
                o[a] = mine ? 1.f : 0.f;
                o[81 + a] = (occ && !mine) ? 1.f : 0.f;
                o[162 + a] = lg ? 1.f : 0.f;
                o[243 + a] = 1.f;               // canonical frame: side == 1 (utac_game.py:106)
            }
        }
    }
}

// one thread per game; same specification as oracle/utac_rules.c:utac_playout
__global__ void k_rules_playouts(uint64_t seed, uint64_t first_game, int n, int max_plies,
                                 tc_playout_result *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t game = first_game + (uint64_t)i;
    Pos p;
    p.mine[0] = p.mine[1] = p.mine[2] = 0;
    p.occ[0] = p.occ[1] = p.occ[2] = 0;
    p.last = 0;
    uint64_t sum = mix64(seed ^ game);
    int ply = 0, outcome = 0;
    int first_to_move = 1;
    Summary s = summarize(p);
    while (ply < max_plies && status_of(s) == 0) {
        uint32_t lw[3], am[3];
        legal_words(p, s, lw);
        to_action_order(lw, am);
        int nl = popc32(am[0]) + popc32(am[1]) + popc32(am[2]);
        uint64_t r = mix64(seed ^ (game * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)ply * 0xD1B54A32D192ED03ull));
        int a = nth_action(am, (int)(r % (uint64_t)nl));
        p = play(p, a);
        s = summarize(p);
        first_to_move = -first_to_move;
        ++ply;
        uint32_t w[8];
        pack(p, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum = mix64(sum ^ w[k]);
        legal_words(p, s, lw);
        to_action_order(lw, am);
#pragma unroll
        for (int k = 0; k < 3; ++k) sum = mix64(sum ^ am[k]);
        int st = status_of(s);
        if (st == 3) outcome = first_to_move;
        if (st == 1) outcome = -first_to_move;
        sum = mix64(sum ^ (uint32_t)st);
    }
    tc_playout_result res;
    res.checksum = sum;
    res.plies = ply;
    res.outcome = outcome;
    uint32_t w[8];
    pack(p, w);
#pragma unroll
    for (int k = 0; k < 8; ++k) res.final_state[k] = w[k];
    out[i] = res;
}

}  // namespace tc

using namespace tc;

namespace {
// Per-thread device scratch of the stateless rules hooks, grown on demand and kept: a hook called once per ply must not
// pay five cudaMalloc / cudaFree pairs per call.  One arena per host thread and device, carved into 256-byte aligned parts.
struct Scratch {
    int device = -1;
    uint8_t *base = nullptr;
    size_t bytes = 0;
    ~Scratch() { if (base) cudaFree(base); }
    cudaError_t reserve(size_t need)
    {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        if (dev == device && need <= bytes) return cudaSuccess;
        if (base) { cudaFree(base); base = nullptr; bytes = 0; }
        device = dev;
        need = (need + (1u << 20)) & ~(size_t)((1u << 20) - 1);
        err = cudaMalloc((void **)&base, need);
        if (err == cudaSuccess) bytes = need;
        return err;
    }
};
thread_local Scratch g_scratch;

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" int tc_rules_step(int n, const uint32_t *states_host, const int8_t *actions_host,
                             uint32_t *out_states_host, int8_t *status_host)
{
    TC_CHECK(n >= 0 && (n == 0 || (states_host && actions_host && out_states_host && status_host)), TC_EINVAL,
             "tc_rules_step: null buffer");
    if (n == 0) return TC_OK;
    const size_t sb = align256(32 * (size_t)n), ab = align256((size_t)n);
    TC_CUDA(g_scratch.reserve(2 * sb + 2 * ab + 256));
    uint32_t *in = reinterpret_cast<uint32_t *>(g_scratch.base), *out = reinterpret_cast<uint32_t *>(g_scratch.base + sb);
    int8_t *act = reinterpret_cast<int8_t *>(g_scratch.base + 2 * sb), *st = act + ab;
    int *bad = reinterpret_cast<int *>(g_scratch.base + 2 * sb + 2 * ab);
    TC_CUDA(cudaMemcpy(in, states_host, 32 * (size_t)n, cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemcpy(act, actions_host, (size_t)n, cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemset(bad, 0, sizeof(int)));
    k_rules_step<<<ceil_div(n, 256), 256>>>(n, in, act, out, st, bad);
    TC_CUDA(cudaGetLastError());
    int nbad = 0;
    TC_CUDA(cudaMemcpy(out_states_host, out, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemcpy(status_host, st, (size_t)n, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemcpy(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
    TC_CHECK(nbad == 0, TC_EILLEGAL, "tc_rules_step: %d illegal move(s) requested", nbad);
    return TC_OK;
}

static int mask_obs_common(int n, const uint32_t *states_host, uint8_t *mask_host, float *obs_host,
                           int8_t *status_host)
{
    TC_CHECK(n >= 0 && (n == 0 || states_host), TC_EINVAL, "rules hook: null states");
    if (n == 0) return TC_OK;
    const size_t sb = align256(32 * (size_t)n), mb = mask_host ? align256(81 * (size_t)n) : 0,
                 ob = obs_host ? align256(324 * sizeof(float) * (size_t)n) : 0, tb = status_host ? align256((size_t)n) : 0;
    TC_CUDA(g_scratch.reserve(sb + mb + ob + tb));
    uint32_t *in = reinterpret_cast<uint32_t *>(g_scratch.base);
    uint8_t *mask = mask_host ? g_scratch.base + sb : nullptr;
    float *obs = obs_host ? reinterpret_cast<float *>(g_scratch.base + sb + mb) : nullptr;
    int8_t *st = status_host ? reinterpret_cast<int8_t *>(g_scratch.base + sb + mb + ob) : nullptr;
    TC_CUDA(cudaMemcpy(in, states_host, 32 * (size_t)n, cudaMemcpyHostToDevice));
    int threads = 256, warps_per_block = threads / 32;
    k_rules_mask_obs<<<ceil_div(n, warps_per_block), threads>>>(n, in, mask, obs, st);
    TC_CUDA(cudaGetLastError());
    if (mask_host) TC_CUDA(cudaMemcpy(mask_host, mask, 81 * (size_t)n, cudaMemcpyDeviceToHost));
    if (obs_host) TC_CUDA(cudaMemcpy(obs_host, obs, 324 * sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    if (status_host) TC_CUDA(cudaMemcpy(status_host, st, (size_t)n, cudaMemcpyDeviceToHost));
    return TC_OK;
}

extern "C" int tc_rules_legal_mask(int n, const uint32_t *states_host, uint8_t *mask_host)
{
    TC_CHECK(n == 0 || mask_host, TC_EINVAL, "tc_rules_legal_mask: null output");
    return mask_obs_common(n, states_host, mask_host, nullptr, nullptr);
}

extern "C" int tc_rules_status(int n, const uint32_t *states_host, int8_t *status_host)
{
    TC_CHECK(n == 0 || status_host, TC_EINVAL, "tc_rules_status: null output");
    return mask_obs_common(n, states_host, nullptr, nullptr, status_host);
}

extern "C" int tc_rules_encode_obs(int n, const uint32_t *states_host, float *obs_host)
{
    TC_CHECK(n == 0 || obs_host, TC_EINVAL, "tc_rules_encode_obs: null output");
    return mask_obs_common(n, states_host, nullptr, obs_host, nullptr);
}

extern "C" int tc_rules_playouts(uint64_t seed, uint64_t first_game, int n, int max_plies,
                                 tc_playout_result *out_host)
{
    TC_CHECK(n >= 0 && (n == 0 || out_host), TC_EINVAL, "tc_rules_playouts: null output");
    if (n == 0) return TC_OK;
    TC_CUDA(g_scratch.reserve(sizeof(tc_playout_result) * (size_t)n));
    tc_playout_result *out = reinterpret_cast<tc_playout_result *>(g_scratch.base);
    k_rules_playouts<<<ceil_div(n, 128), 128>>>(seed, first_game, n, max_plies, out);
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaMemcpy(out_host, out, sizeof(tc_playout_result) * (size_t)n, cudaMemcpyDeviceToHost));
    return TC_OK;
}
