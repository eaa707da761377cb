// Warp-cooperative forms of the rules (rules.cuh) for kernels that own a position per warp: the tree kernel's leaf
// expansion and the trunk kernels' plane encoding.  Same results as the scalar functions the rules hooks and the host
// GameState use - every search-parity test goes through them.
#pragma once
#include "rules.cuh"

namespace tc {

// one completed line among the 9 bits of a 3x3 grid (the same four tests as rules.cuh:line_fields, one field)
__device__ __forceinline__ bool line9(uint32_t b)
{
    return ((b & (b >> 1) & (b >> 2) & 0x049u) | (b & (b >> 3) & (b >> 6) & 0x007u) | (b & (b >> 4) & (b >> 8) & 1u) |
            ((b >> 2) & (b >> 4) & (b >> 6) & 1u)) != 0u;
}

// Action-order legal mask and terminal status of a position, computed BY the warp instead of 32 times over: lane i < 9
// owns sub-board i (line / full tests on its 9 bits, three ballots give the won / closed sets of rules.cuh:summarize),
// then lane l tests its actions l, l + 32, l + 64 (rules.cuh:legal_words + action_bit).  With MASKS the stones of the
// side to move and all stones come back in action order as well (the observation planes).
template <bool MASKS = false>
__device__ __forceinline__ void warp_legal_mask(const Pos &p, int lane, uint32_t am[3], int &status, uint32_t *mine_am = nullptr,
                                                uint32_t *occ_am = nullptr)
{
    constexpr uint32_t ALL = 0xFFFFFFFFu;
    const int i = lane < 9 ? lane : 0;
    const int g = (i * 11) >> 5;                                  // i / 3
    const int sh = 9 * (i - 3 * g);
    const uint32_t wm = g == 0 ? p.mine[0] : (g == 1 ? p.mine[1] : p.mine[2]);
    const uint32_t wo = g == 0 ? p.occ[0] : (g == 1 ? p.occ[1] : p.occ[2]);
    const uint32_t mine = (wm >> sh) & F9, occ = (wo >> sh) & F9;
    const bool sub = lane < 9;
    const uint32_t won_mine = __ballot_sync(ALL, sub && line9(mine));
    const uint32_t won_opp = __ballot_sync(ALL, sub && line9(occ & ~mine));
    const uint32_t closed = won_mine | won_opp | __ballot_sync(ALL, sub && occ == F9);
    status = line9(won_opp) ? 1 : (line9(won_mine) ? 3 : (closed == F9 ? 2 : 0));        // rules.cuh:status_of
    uint32_t open = ~closed & F9;
    if (p.last != 0) {
        const uint32_t forced = 1u << (p.last - 1);
        if (open & forced) open = forced;                         // forced sub-board still open; otherwise free move
    }
    if (status != 0) open = 0u;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int a = lane + 32 * q;
        const int r = (a * 57) >> 9, c = a - 9 * r;               // a / 9, a % 9 (exact for a < 96)
        const int r3 = (r * 11) >> 5, c3 = (c * 11) >> 5;         // / 3 (exact below 11)
        const int cell = 3 * (r - 3 * r3) + (c - 3 * c3);
        const uint32_t ow = r3 == 0 ? p.occ[0] : (r3 == 1 ? p.occ[1] : p.occ[2]);
        const bool stone = a < 81 && ((ow >> (9 * c3 + cell)) & 1u);
        const bool lg = a < 81 && ((open >> (3 * r3 + c3)) & 1u) && !stone;
        am[q] = __ballot_sync(ALL, lg);
        if (MASKS) {
            const uint32_t mw = r3 == 0 ? p.mine[0] : (r3 == 1 ? p.mine[1] : p.mine[2]);
            occ_am[q] = __ballot_sync(ALL, stone);
            mine_am[q] = __ballot_sync(ALL, a < 81 && ((mw >> (9 * c3 + cell)) & 1u));
        }
    }
}

}  // namespace tc
