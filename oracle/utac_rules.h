/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C restatement of the Ultimate Tic-Tac-Toe rules engine the reference
 * calls through `utac_gym.core.GameState` (external C++ library `utac`, wrapped
 * with nanobind; un-vendored and un-pinned: /root/reference/README.md:14-15,150-152,
 * absent from /root/reference/pyproject.toml:13-20).
 *
 * PARITY UNPINNED: the reference tree holds neither the utac sources nor any
 * golden vector / known-answer test for it.  This file
 * restates *standard* UTTT rules under the layout facts the reference's own
 * call sites pin down:
 *   (1) bit k of board[i]/occ[i] <-> global cell (3*(i/3)+k/3, 3*(i%3)+k%3)
 *       /root/reference/src/tacult/utac_game.py:259-265,283-291
 *   (2) board[i] is a subset of occ[i]; colour flip = ~board & occ
 *       /root/reference/src/tacult/utac_game.py:100-106
 *   (3) main_board subset of main_occ, same flip       (utac_game.py:104)
 *   (4) last_square in {-1, 0..8}, -1 only at game start (utac_game.py:182-184,215-216)
 *   (5) search key = board, occ, side, last_square only (utac_game.py:222-241)
 *   (6) action index is global row-major 0..80         (utac_game.py:130,171)
 *   (7) terminal_value() in {0 draw, +-1 winner}, first mover is +1
 *       (utac_game.py:64-81,243-244; base/coach.py:90)
 *   (8) get_obs() -> 324 values reshaped (4,9,9)        (utac_game.py:247-250)
 *   (9) draws exist (scored 1e-4)                       (utac_game.py:78-79)
 * Rule choices the reference does NOT pin are resolved as:
 *   - a sub-board is closed once won or full; closed sub-boards accept no moves;
 *   - a move into cell k sends the opponent to sub-board k; if that sub-board
 *     is closed the opponent may play any empty cell of any open sub-board;
 *   - three won sub-boards in a line win the game; otherwise all nine closed
 *     is a draw;
 *   - observation planes: 0 = `board` stones, 1 = occ & ~board, 2 = legal-move
 *     cells, 3 = all ones iff side == 1.
 *
 * Written deliberately as table-free, loop-over-lines scalar code so that it
 * shares no structure with the bit-parallel CUDA implementation it checks.
 */
#ifndef ORACLE_UTAC_RULES_H
#define ORACLE_UTAC_RULES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint16_t board[9];    /* stones of the player whose id is +1 (9 bits per sub-board) */
    uint16_t occ[9];      /* all stones */
    uint16_t main_board;  /* sub-boards won by player +1 */
    uint16_t main_occ;    /* sub-boards won by either player */
    uint16_t game_occ;    /* sub-boards closed (won or full) */
    int8_t   side;        /* +1 or -1: player to move */
    int8_t   last_square; /* cell index of the last move, -1 at game start */
} utac_state;

void utac_init(utac_state *s);
/* recompute main_board/main_occ/game_occ from board/occ (fact 5: they are derivable) */
void utac_rederive(utac_state *s);
int  utac_is_legal(const utac_state *s, int action);
/* returns 0 on success, -1 if the move is illegal (state untouched) */
int  utac_make_move(utac_state *s, int action);
/* writes 81 0/1 bytes in action order; returns the number of legal moves */
int  utac_valid_mask(const utac_state *s, uint8_t out[81]);
int  utac_is_terminal(const utac_state *s);
/* 0 = draw or not finished; +1 / -1 = winner id */
int  utac_terminal_value(const utac_state *s);
/* 4x9x9 float planes */
void utac_obs(const utac_state *s, float out[324]);
/* colour flip used by UtacGame.getCanonicalForm (utac_game.py:97-107) */
void utac_flip(utac_state *s);

/* ---- deterministic playout driver shared (by specification) with the CUDA side ---- */
/* 64-bit mix (splitmix64 finaliser) of seed, game index, ply */
uint64_t utac_mix64(uint64_t x);
/*
 * Plays game `game` from the initial position choosing, at every ply, the
 * (r % n_legal)-th legal move in action order with r = utac_mix64(seed ^ game*K1 ^ ply*K2).
 * After every ply the canonical successor (side-to-move stones in `board`) is
 * folded into a running checksum together with its legal mask and terminal
 * status.  Outputs: number of plies, outcome (+1/-1 first/second player won,
 * 0 draw), final canonical packed state words and the checksum.
 */
typedef struct {
    uint64_t checksum;
    uint32_t final_state[8];
    int32_t  plies;
    int32_t  outcome;
} utac_playout_result;
void utac_playout(uint64_t seed, uint64_t game, int max_plies, utac_playout_result *out);
void utac_playout_many(uint64_t seed, uint64_t first_game, int n, int max_plies, utac_playout_result *out);

/* canonical device packing (see include/tacult_b200.h): 8 x u32 */
void utac_pack_canonical(const utac_state *s, uint32_t out[8]);
void utac_unpack_canonical(const uint32_t in[8], utac_state *s);

#ifdef __cplusplus
}
#endif
#endif
