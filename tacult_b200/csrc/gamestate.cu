// Host-side game-state object of the C ABI (tc_gs_*): the value-semantic `utac_gym.core.GameState` the reference's
// adapter and callers hold on the HOST (/root/reference/src/tacult/utac_game.py:18,45-46,62,75-81,100-107,240-248;
// base/coach.py:17-18,68,234).  In the reference this object lives in the external C++ `utac` library; here it is the
// product's own bit-parallel rules (rules.cuh) compiled for the host.  It is not a fallback for anything on the search
// path: the trees, leaves and planes of the search never touch it.  Its only throughput role is tc_gs_pack, which
// turns the caller's list of GameState objects into the packed canonical roots of tc_set_roots_packed in one call.
#include <string.h>

#include "common.cuh"
#include "rules.cuh"

using namespace tc;

namespace {

// position as seen by the side to move (the frame rules.cuh works in)
inline Pos mover_frame(const tc_gamestate *s)
{
    Pos p;
    for (int g = 0; g < 3; ++g) {
        uint32_t b = 0, o = 0;
        for (int j = 0; j < 3; ++j) {
            b |= ((uint32_t)s->board[3 * g + j] & F9) << (9 * j);
            o |= ((uint32_t)s->occ[3 * g + j] & F9) << (9 * j);
        }
        p.occ[g] = o;
        p.mine[g] = (s->side == 1) ? (b & o) : (o & ~b);
    }
    p.last = (uint32_t)(s->last_square + 1) & 0xFu;
    return p;
}

inline void derive_fields(tc_gamestate *s)
{
    Pos p = mover_frame(s);
    Summary m = summarize(p);
    const uint32_t plus = s->side == 1 ? m.won_mine : m.won_opp;      // sub-boards won by player +1
    s->main_board = (uint16_t)plus;
    s->main_occ = (uint16_t)(m.won_mine | m.won_opp);
    s->game_occ = (uint16_t)m.closed;
}

}  // namespace

extern "C" void tc_gs_init(tc_gamestate *s)
{
    memset(s, 0, sizeof *s);
    s->side = 1;
    s->last_square = -1;
}

extern "C" void tc_gs_rederive(tc_gamestate *s) { derive_fields(s); }

extern "C" int tc_gs_valid_mask(const tc_gamestate *s, double *mask81)
{
    Pos p = mover_frame(s);
    uint32_t lw[3], am[3];
    legal_words(p, summarize(p), lw);
    to_action_order(lw, am);
    int n = 0;
    for (int a = 0; a < 81; ++a) {
        const bool on = (am[a >> 5] >> (a & 31)) & 1u;
        if (mask81) mask81[a] = on ? 1.0 : 0.0;
        n += on;
    }
    return n;
}

extern "C" int tc_gs_make_move(tc_gamestate *s, int action)
{
    Pos p = mover_frame(s);
    uint32_t lw[3];
    legal_words(p, summarize(p), lw);
    TC_CHECK(action >= 0 && action <= 80 && action_bit(lw, action), TC_EILLEGAL, "tc_gs_make_move: illegal move %d",
             action);
    int word, bit, cell;
    action_slot(action, word, bit, cell);
    const int sub = 3 * word + bit / 9;
    s->occ[sub] |= (uint16_t)(1u << cell);
    if (s->side == 1) s->board[sub] |= (uint16_t)(1u << cell);
    s->side = (int8_t)-s->side;
    s->last_square = (int8_t)cell;
    derive_fields(s);
    return TC_OK;
}

extern "C" int tc_gs_is_terminal(const tc_gamestate *s) { return status_of(summarize(mover_frame(s))) != 0; }

// 0: draw or game in progress; +1 / -1: the winner's id (the frame of current_player(), utac_game.py:78-81)
extern "C" int tc_gs_terminal_value(const tc_gamestate *s)
{
    const int st = status_of(summarize(mover_frame(s)));
    if (st == 3) return s->side;
    if (st == 1) return -s->side;
    return 0;
}

// planes (utac_game.py:247-250, README.md:97-105): `board` stones, the other stones, legal cells, side == 1
extern "C" void tc_gs_obs(const tc_gamestate *s, float *obs324)
{
    Pos p = mover_frame(s);
    uint32_t lw[3];
    legal_words(p, summarize(p), lw);
    uint32_t plus[3];
    for (int g = 0; g < 3; ++g) plus[g] = (s->side == 1) ? p.mine[g] : (p.occ[g] & ~p.mine[g]);
    for (int a = 0; a < 81; ++a) {
        const bool b = action_bit(plus, a), o = action_bit(p.occ, a);
        obs324[a] = b ? 1.f : 0.f;
        obs324[81 + a] = (o && !b) ? 1.f : 0.f;
        obs324[162 + a] = action_bit(lw, a) ? 1.f : 0.f;
        obs324[243 + a] = s->side == 1 ? 1.f : 0.f;
    }
}

// canonicalBoards of getActionProbs (mcts.py:175) -> packed roots.  active[i] == 0 on entry marks a None board.
// A board whose side to move is not +1 is accepted only when the game is over (the arena keeps passing finished
// boards, arena.py:176-184; the reference's search merely looks their terminal value up): it is marked inactive.
extern "C" int tc_gs_pack(int n, const tc_gamestate *gs_host, uint32_t *states_host, uint8_t *active_host)
{
    TC_CHECK(n >= 0 && (n == 0 || (gs_host && states_host && active_host)), TC_EINVAL, "tc_gs_pack: null buffer");
    for (int i = 0; i < n; ++i) {
        uint32_t *w = states_host + 8 * (size_t)i;
        if (!active_host[i]) {
            memset(w, 0, 8 * sizeof(uint32_t));
            continue;
        }
        const tc_gamestate *s = gs_host + i;
        if (s->side != 1) {
            TC_CHECK(tc_gs_is_terminal(s), TC_EINVAL,
                     "tc_gs_pack: board %d is not in canonical form (side == %d); pass getCanonicalForm's result", i,
                     (int)s->side);
            active_host[i] = 0;
            memset(w, 0, 8 * sizeof(uint32_t));
            continue;
        }
        pack(mover_frame(s), w);
    }
    return TC_OK;
}

// Move choice from visit counts (coach.py:125-140, mcts.py:200-214): see the header.  Equals, draw for draw, the numpy
// formulation  cum = cumsum(counts); a = argmax(cum > floor(u * cum[-1]))  /  counts.argmax()  that bench.py used before.
extern "C" int tc_choose_actions(int n, const uint32_t *counts_host, const uint8_t *tau_one_host, const double *u_host,
                                 int8_t *actions_host)
{
    TC_CHECK(n >= 0 && (n == 0 || (counts_host && tau_one_host && actions_host)), TC_EINVAL, "tc_choose_actions: null buffer");
    for (int i = 0; i < n; ++i) {
        const uint32_t *c = counts_host + 81 * (size_t)i;
        int best = 0;
        if (tau_one_host[i]) {
            TC_CHECK(u_host != nullptr, TC_EINVAL, "tc_choose_actions: tau = 1 rows need uniform draws");
            int64_t total = 0;
            for (int a = 0; a < 81; ++a) total += c[a];
            const int64_t draw = (int64_t)(u_host[i] * (double)total);
            int64_t cum = 0;
            for (int a = 0; a < 81; ++a) {
                cum += c[a];
                if (cum > draw) { best = a; break; }
            }
        } else {
            uint32_t top = c[0];
            for (int a = 1; a < 81; ++a)
                if (c[a] > top) { top = c[a]; best = a; }
        }
        actions_host[i] = (int8_t)best;
    }
    return TC_OK;
}
