/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see utac_rules.h for scope, citations and
 * the "PARITY UNPINNED" note).  Scalar, loop-based UTTT rules.
 */
#include "utac_rules.h"
#include <string.h>

/* the eight lines of a 3x3 grid, as cell triples */
static const int LINES[8][3] = {
    {0, 1, 2}, {3, 4, 5}, {6, 7, 8},
    {0, 3, 6}, {1, 4, 7}, {2, 5, 8},
    {0, 4, 8}, {2, 4, 6},
};

static int has_line(unsigned bits)
{
    for (int l = 0; l < 8; ++l) {
        int hit = 1;
        for (int j = 0; j < 3; ++j)
            if (!((bits >> LINES[l][j]) & 1u)) hit = 0;
        if (hit) return 1;
    }
    return 0;
}

/* action (global row-major, fact 6) -> sub-board and cell (fact 1) */
static void split_action(int action, int *sub, int *cell)
{
    int r = action / 9, c = action % 9;
    *sub = (r / 3) * 3 + (c / 3);
    *cell = (r % 3) * 3 + (c % 3);
}

void utac_init(utac_state *s)
{
    memset(s, 0, sizeof *s);
    s->side = 1;
    s->last_square = -1;
}

void utac_rederive(utac_state *s)
{
    s->main_board = s->main_occ = s->game_occ = 0;
    for (int i = 0; i < 9; ++i) {
        unsigned mine = s->board[i] & 0x1FFu;
        unsigned theirs = (s->occ[i] & ~s->board[i]) & 0x1FFu;
        int w1 = has_line(mine), w2 = has_line(theirs);
        if (w1) s->main_board |= (uint16_t)(1u << i);
        if (w1 || w2) s->main_occ |= (uint16_t)(1u << i);
        if (w1 || w2 || (s->occ[i] & 0x1FFu) == 0x1FFu) s->game_occ |= (uint16_t)(1u << i);
    }
}

int utac_is_terminal(const utac_state *s)
{
    unsigned p1 = s->main_board & 0x1FFu;
    unsigned p2 = (s->main_occ & ~s->main_board) & 0x1FFu;
    if (has_line(p1) || has_line(p2)) return 1;
    return (s->game_occ & 0x1FFu) == 0x1FFu;
}

int utac_terminal_value(const utac_state *s)
{
    unsigned p1 = s->main_board & 0x1FFu;
    unsigned p2 = (s->main_occ & ~s->main_board) & 0x1FFu;
    if (has_line(p1)) return 1;
    if (has_line(p2)) return -1;
    return 0;
}

/* legality of one action in a position already known not to be terminal */
static int legal_in_live_position(const utac_state *s, int action)
{
    int sub, cell;
    split_action(action, &sub, &cell);
    if ((s->occ[sub] >> cell) & 1u) return 0;          /* cell taken */
    if ((s->game_occ >> sub) & 1u) return 0;           /* sub-board closed */
    if (s->last_square >= 0 && !((s->game_occ >> s->last_square) & 1u))
        return sub == s->last_square;                  /* forced sub-board still open */
    return 1;                                          /* free move */
}

int utac_is_legal(const utac_state *s, int action)
{
    if (action < 0 || action > 80) return 0;
    if (utac_is_terminal(s)) return 0;
    return legal_in_live_position(s, action);
}

int utac_valid_mask(const utac_state *s, uint8_t out[81])
{
    int n = 0;
    int live = !utac_is_terminal(s);
    for (int a = 0; a < 81; ++a) {
        out[a] = (uint8_t)(live && legal_in_live_position(s, a));
        n += out[a];
    }
    return n;
}

int utac_make_move(utac_state *s, int action)
{
    if (!utac_is_legal(s, action)) return -1;
    int sub, cell;
    split_action(action, &sub, &cell);
    s->occ[sub] |= (uint16_t)(1u << cell);
    if (s->side == 1) s->board[sub] |= (uint16_t)(1u << cell);
    unsigned mover_bits = (s->side == 1) ? s->board[sub] : (uint16_t)(s->occ[sub] & ~s->board[sub]);
    if (has_line(mover_bits & 0x1FFu)) {
        s->main_occ |= (uint16_t)(1u << sub);
        if (s->side == 1) s->main_board |= (uint16_t)(1u << sub);
        s->game_occ |= (uint16_t)(1u << sub);
    } else if ((s->occ[sub] & 0x1FFu) == 0x1FFu) {
        s->game_occ |= (uint16_t)(1u << sub);
    }
    s->side = (int8_t)(-s->side);
    s->last_square = (int8_t)cell;
    return 0;
}

void utac_obs(const utac_state *s, float out[324])
{
    uint8_t legal[81];
    utac_valid_mask(s, legal);
    for (int a = 0; a < 81; ++a) {
        int sub, cell;
        split_action(a, &sub, &cell);
        unsigned b = (s->board[sub] >> cell) & 1u;
        unsigned o = (s->occ[sub] >> cell) & 1u;
        out[a] = (float)b;
        out[81 + a] = (float)(o & ~b & 1u);
        out[162 + a] = (float)legal[a];
        out[243 + a] = (s->side == 1) ? 1.0f : 0.0f;
    }
}

void utac_flip(utac_state *s)
{
    for (int i = 0; i < 9; ++i)
        s->board[i] = (uint16_t)(~s->board[i] & s->occ[i]);
    s->main_board = (uint16_t)(~s->main_board & s->main_occ);
    s->side = 1;
}

/* ------------------------------------------------------------------------- */

void utac_pack_canonical(const utac_state *s, uint32_t out[8])
{
    for (int g = 0; g < 3; ++g) {
        out[g] = 0;
        out[3 + g] = 0;
        for (int j = 0; j < 3; ++j) {
            out[g] |= (uint32_t)(s->board[3 * g + j] & 0x1FFu) << (9 * j);
            out[3 + g] |= (uint32_t)(s->occ[3 * g + j] & 0x1FFu) << (9 * j);
        }
    }
    out[6] = (uint32_t)(s->last_square + 1);
    out[7] = 0;
}

void utac_unpack_canonical(const uint32_t in[8], utac_state *s)
{
    memset(s, 0, sizeof *s);
    for (int i = 0; i < 9; ++i) {
        s->board[i] = (uint16_t)((in[i / 3] >> (9 * (i % 3))) & 0x1FFu);
        s->occ[i] = (uint16_t)((in[3 + i / 3] >> (9 * (i % 3))) & 0x1FFu);
    }
    s->side = 1;
    s->last_square = (int8_t)((int)(in[6] & 0xFu) - 1);
    utac_rederive(s);
}

uint64_t utac_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void utac_playout(uint64_t seed, uint64_t game, int max_plies, utac_playout_result *out)
{
    utac_state s;
    utac_init(&s);
    uint64_t sum = utac_mix64(seed ^ game);
    int ply = 0;
    int first_player_to_move = 1;   /* +1 while the first player is to move */
    int outcome = 0;
    while (ply < max_plies && !utac_is_terminal(&s)) {
        uint8_t legal[81];
        int n = utac_valid_mask(&s, legal);
        uint64_t r = utac_mix64(seed ^ (game * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)ply * 0xD1B54A32D192ED03ull));
        int k = (int)(r % (uint64_t)n);
        int action = -1;
        for (int a = 0; a < 81; ++a)
            if (legal[a] && k-- == 0) { action = a; break; }
        utac_make_move(&s, action);
        /* canonical successor: the side to move owns `board` (utac_game.py:97-107) */
        if (s.side != 1) utac_flip(&s);
        first_player_to_move = -first_player_to_move;
        ++ply;

        uint32_t w[8];
        utac_pack_canonical(&s, w);
        for (int i = 0; i < 8; ++i) sum = utac_mix64(sum ^ w[i]);
        uint32_t m[3] = {0, 0, 0};
        utac_valid_mask(&s, legal);
        for (int a = 0; a < 81; ++a)
            if (legal[a]) m[a >> 5] |= 1u << (a & 31);
        for (int i = 0; i < 3; ++i) sum = utac_mix64(sum ^ m[i]);
        uint32_t status = 0;
        if (utac_is_terminal(&s)) {
            int tv = utac_terminal_value(&s);   /* in the canonical frame +1 = side to move */
            status = (tv == 0) ? 2u : (tv == 1 ? 3u : 1u);
            if (tv != 0) outcome = (tv == 1) ? first_player_to_move : -first_player_to_move;
        }
        sum = utac_mix64(sum ^ status);
    }
    out->checksum = sum;
    out->plies = ply;
    out->outcome = outcome;
    utac_pack_canonical(&s, out->final_state);
}

void utac_playout_many(uint64_t seed, uint64_t first_game, int n, int max_plies, utac_playout_result *out)
{
    for (int i = 0; i < n; ++i)
        utac_playout(seed, first_game + (uint64_t)i, max_plies, &out[i]);
}
