"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's game adapter `UtacGame`
(/root/reference/src/tacult/utac_game.py:8-250) over the oracle rules engine
(oracle/utac_gym, PARITY UNPINNED — see oracle/utac_rules.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from utac_gym.core import GameState  # noqa: E402

ACTIONS = 81
DRAW_VALUE = 1e-4  # utac_game.py:78-79

# action a (global row-major) -> (sub-board, cell): utac_game.py:259-291
_R, _C = np.divmod(np.arange(81), 9)
ACTION_SUB = (_R // 3) * 3 + _C // 3
ACTION_CELL = (_R % 3) * 3 + _C % 3


class OracleGame:
    """Same protocol as the reference's `Game` (base/game.py:1-113)."""

    def getInitBoard(self):                      # utac_game.py:12-18
        return GameState()

    def getBoardSize(self):                      # utac_game.py:20-25
        return (9, 9, 4)

    def getActionSize(self):                     # utac_game.py:27-32
        return ACTIONS

    def getNextState(self, board, player, action):   # utac_game.py:34-47
        nxt = GameState(board)
        nxt.make_move(action)
        return nxt, -player

    def getValidMoves(self, board, player):      # utac_game.py:49-62
        mover = 1 if board.current_player() == 1 else -1
        if player != mover:
            return np.zeros(ACTIONS)
        return board.get_valid_mask()

    def getGameEnded(self, board, player):       # utac_game.py:64-81
        if not board.is_terminal():
            return 0
        tv = board.terminal_value()
        if tv == 0:
            return DRAW_VALUE
        return 1 if tv == player else -1

    def getCanonicalForm(self, board, player):   # utac_game.py:83-107
        if player == 1:
            return board
        gs = GameState(board)._get_gs()
        gs.board = (~gs.board) & gs.occ
        gs.main_board = (~gs.main_board) & gs.main_occ
        gs.side = 1
        return GameState(gs)

    def stringRepresentation(self, board):       # utac_game.py:222-241
        gs = board._get_gs()
        left = "".join(np.binary_repr(int(x), width=9) for x in gs.board)
        right = "".join(np.binary_repr(int(x), width=9) for x in gs.occ)
        return left + right + str(gs.side) + str(gs.last_square)

    def _get_obs(self, board):                   # utac_game.py:247-250
        import torch
        return torch.from_numpy(np.array(board.get_obs()).reshape(4, 9, 9)).float()

    # -- the 8 dihedral images of (board, pi): utac_game.py:109-220 ---------
    def getSymmetries(self, board, pi):
        gs = board._get_gs()
        grid_mine = _bitboards_to_grid(gs.board)
        grid_occ = _bitboards_to_grid(gs.occ)
        g_closed = _bits9(gs.game_occ).reshape(3, 3)
        g_won = _bits9(gs.main_occ).reshape(3, 3)
        g_mine = _bits9(gs.main_board).reshape(3, 3)
        last = np.zeros(9, dtype=bool)
        if gs.last_square != -1:
            last[gs.last_square] = True
        last = last.reshape(3, 3)
        pi_grid = np.reshape(pi, (9, 9))
        out = []
        for quarter_turns in range(4):
            for mirror in (False, True):
                def tf(x):
                    y = np.rot90(x, quarter_turns)
                    return np.fliplr(y) if mirror else y
                img = board._get_gs()
                img.occ = _grid_to_bitboards(tf(grid_occ))
                img.board = _grid_to_bitboards(tf(grid_mine))
                img.game_occ = _pack9(tf(g_closed).ravel())
                img.main_occ = _pack9(tf(g_won).ravel())
                img.main_board = _pack9(tf(g_mine).ravel())
                if gs.last_square != -1:
                    img.last_square = int(np.argmax(tf(last).ravel()))
                out.append((GameState(img), tf(pi_grid).flatten()))
        return out


def _bits9(x):
    return np.array([(int(x) >> k) & 1 for k in range(9)], dtype=np.int8)


def _pack9(bits):
    return int(sum(int(bool(b)) << k for k, b in enumerate(bits)))


def _bitboards_to_grid(bbs):
    grid = np.zeros((9, 9), dtype=np.int8)
    for i, bb in enumerate(bbs):
        grid[(i // 3) * 3:(i // 3) * 3 + 3, (i % 3) * 3:(i % 3) * 3 + 3] = _bits9(bb).reshape(3, 3)
    return grid


def _grid_to_bitboards(grid):
    return np.array([_pack9(grid[(i // 3) * 3:(i // 3) * 3 + 3, (i % 3) * 3:(i % 3) * 3 + 3].ravel())
                     for i in range(9)], dtype=np.uint16)
