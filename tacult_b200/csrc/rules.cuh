// Bit-parallel Ultimate Tic-Tac-Toe rules on the packed canonical position.
//
// Replaces the external C++ `utac` engine the reference reaches through utac_gym
// (call sites /root/reference/src/tacult/utac_game.py:45-46,62,75-81,100-107,248).
// Rule choices follow SURVEY.md §10.3 (the reference does not pin them).
//
// Layout: three 27-bit words per colour set; sub-board i lives in word i/3 at bit offset
// 9*(i%3); cell k of a sub-board is bit k, i.e. global cell (3*(i/3)+k/3, 3*(i%3)+k%3)
// (utac_game.py:259-265).  All eight line tests of three sub-boards run at once on one word.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TC_HD __host__ __device__ __forceinline__
#else
#define TC_HD inline
#endif

namespace tc {

constexpr uint32_t F9 = 0x1FFu;                                  // one sub-board
constexpr uint32_t REP3 = 1u | (1u << 9) | (1u << 18);           // replicate a 9-bit pattern
constexpr uint32_t ROW_SEL = 0x049u * REP3;                      // cells 0,3,6 of each field
constexpr uint32_t COL_SEL = 0x007u * REP3;                      // cells 0,1,2
constexpr uint32_t ONE_SEL = 0x001u * REP3;                      // cell 0
constexpr uint32_t ALL27 = F9 * REP3;

struct Pos {            // canonical position: `mine` = stones of the side to move
    uint32_t mine[3];
    uint32_t occ[3];
    uint32_t last;      // last_square + 1 (0 = game start)
};

struct Summary {        // per-sub-board facts, one bit per sub-board (bit i = sub-board i)
    uint32_t won_mine;
    uint32_t won_opp;
    uint32_t closed;    // won or full
};

// bit j (j=0..2) set iff 9-bit field j of x holds a completed line
TC_HD uint32_t line_fields(uint32_t x)
{
    uint32_t t = (x & (x >> 1) & (x >> 2) & ROW_SEL)
               | (x & (x >> 3) & (x >> 6) & COL_SEL)
               | (x & (x >> 4) & (x >> 8) & ONE_SEL)
               | ((x >> 2) & (x >> 4) & (x >> 6) & ONE_SEL);
    return ((t & F9) ? 1u : 0u) | ((t & (F9 << 9)) ? 2u : 0u) | ((t & (F9 << 18)) ? 4u : 0u);
}

TC_HD uint32_t full_fields(uint32_t x)
{
    return (((x & F9) == F9) ? 1u : 0u) | ((((x >> 9) & F9) == F9) ? 2u : 0u) | ((((x >> 18) & F9) == F9) ? 4u : 0u);
}

TC_HD bool has_line9(uint32_t b) { return (line_fields(b & F9) & 1u) != 0; }

TC_HD Summary summarize(const Pos &p)
{
    Summary s;
    s.won_mine = 0; s.won_opp = 0; s.closed = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        uint32_t wm = line_fields(p.mine[g]);
        uint32_t wo = line_fields(p.occ[g] & ~p.mine[g]);
        uint32_t fu = full_fields(p.occ[g]);
        s.won_mine |= wm << (3 * g);
        s.won_opp |= wo << (3 * g);
        s.closed |= (wm | wo | fu) << (3 * g);
    }
    return s;
}

// TC_LIVE / TC_LOST / TC_DRAW / TC_WON from the side to move's point of view
TC_HD int status_of(const Summary &s)
{
    if (has_line9(s.won_opp)) return 1;
    if (has_line9(s.won_mine)) return 3;
    if (s.closed == F9) return 2;
    return 0;
}

// expand a 3-bit sub-board set (for word g) into a 27-bit cell mask
TC_HD uint32_t spread3(uint32_t bits3)
{
    return ((bits3 & 1u) ? F9 : 0u) | ((bits3 & 2u) ? (F9 << 9) : 0u) | ((bits3 & 4u) ? (F9 << 18) : 0u);
}

// legal cells in board-major packing (same layout as Pos::occ); all zero when the game is over
TC_HD void legal_words(const Pos &p, const Summary &s, uint32_t out[3])
{
    if (status_of(s) != 0) { out[0] = out[1] = out[2] = 0; return; }
    uint32_t open = ~s.closed & F9;
    if (p.last != 0) {
        uint32_t forced = 1u << (p.last - 1);
        if (open & forced) open = forced;       // forced sub-board still open; otherwise free move
    }
#pragma unroll
    for (int g = 0; g < 3; ++g)
        out[g] = spread3((open >> (3 * g)) & 7u) & ~p.occ[g] & ALL27;
}

// action a (global row-major, utac_game.py:130,171) -> word index and bit inside that word
TC_HD void action_slot(int a, int &word, int &bit, int &cell)
{
    int r = a / 9, c = a - 9 * r;
    int rr = r % 3, cc = c % 3;
    word = r / 3;
    cell = 3 * rr + cc;
    bit = 9 * (c / 3) + cell;
}

TC_HD bool action_bit(const uint32_t words[3], int a)
{
    int w, b, cell;
    action_slot(a, w, b, cell);
    return (words[w] >> b) & 1u;
}

// one row (9 cells, global columns 0..8) of a board-major word triple
TC_HD uint32_t row_bits(const uint32_t words[3], int r)
{
    uint32_t w = words[r / 3] >> (3 * (r % 3));
    return (w & 7u) | (((w >> 9) & 7u) << 3) | (((w >> 18) & 7u) << 6);
}

// board-major words -> 81-bit mask in action order, as three 32-bit words
TC_HD void to_action_order(const uint32_t words[3], uint32_t out[3])
{
    uint64_t lo = 0, hi = 0;   // bits 0..62 in lo (7 rows), bits 63..80 in hi
#pragma unroll
    for (int r = 0; r < 7; ++r) lo |= (uint64_t)row_bits(words, r) << (9 * r);
    hi = (uint64_t)row_bits(words, 7) | ((uint64_t)row_bits(words, 8) << 9);
    out[0] = (uint32_t)lo;
    out[1] = (uint32_t)(lo >> 32) | (uint32_t)((hi & 1u) << 31);
    out[2] = (uint32_t)(hi >> 1);
}

// make_move followed by the colour flip of getCanonicalForm (utac_game.py:34-47,97-107):
// the successor as seen by the opponent, who is now to move.
TC_HD Pos play(const Pos &p, int a)
{
    int w, b, cell;
    action_slot(a, w, b, cell);
    Pos n;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        n.mine[g] = p.occ[g] & ~p.mine[g];      // the opponent's stones become `mine`
        n.occ[g] = p.occ[g];
    }
    n.occ[w] |= 1u << b;
    n.last = (uint32_t)cell + 1u;
    return n;
}

TC_HD Pos unpack(const uint32_t w[8])
{
    Pos p;
    p.mine[0] = w[0] & ALL27; p.mine[1] = w[1] & ALL27; p.mine[2] = w[2] & ALL27;
    p.occ[0] = w[3] & ALL27;  p.occ[1] = w[4] & ALL27;  p.occ[2] = w[5] & ALL27;
    p.last = w[6] & 0xFu;
    return p;
}

TC_HD void pack(const Pos &p, uint32_t w[8])
{
    w[0] = p.mine[0]; w[1] = p.mine[1]; w[2] = p.mine[2];
    w[3] = p.occ[0];  w[4] = p.occ[1];  w[5] = p.occ[2];
    w[6] = p.last;    w[7] = 0;
}

TC_HD uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

TC_HD int popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// index (0-based) of the k-th set bit of an action-order mask; k < popcount
TC_HD int nth_action(const uint32_t m[3], int k)
{
    for (int w = 0; w < 3; ++w) {
        int c = popc32(m[w]);
        if (k < c) {
            uint32_t x = m[w];
            for (int i = 0; i < k; ++i) x &= x - 1;
#ifdef __CUDA_ARCH__
            return 32 * w + (__ffs(x) - 1);
#else
            return 32 * w + (__builtin_ffs(x) - 1);
#endif
        }
        k -= c;
    }
    return -1;
}

}  // namespace tc
