"""ORACLE — TEST INFRASTRUCTURE ONLY.

``GAMESTATE``: the raw struct the reference reads and mutates at
/root/reference/src/tacult/utac_game.py:100-106,174-216,240-241.
"""
import numpy as np


class GAMESTATE:
    __slots__ = ("board", "occ", "main_board", "main_occ", "game_occ", "side", "last_square")

    def __init__(self, board=None, occ=None, main_board=0, main_occ=0, game_occ=0, side=1, last_square=-1):
        self.board = np.zeros(9, dtype=np.uint16) if board is None else np.array(board, dtype=np.uint16)
        self.occ = np.zeros(9, dtype=np.uint16) if occ is None else np.array(occ, dtype=np.uint16)
        self.main_board = int(main_board)
        self.main_occ = int(main_occ)
        self.game_occ = int(game_occ)
        self.side = int(side)
        self.last_square = int(last_square)

    def __repr__(self):
        return (f"GAMESTATE(board={self.board.tolist()}, occ={self.occ.tolist()}, main_board={self.main_board}, "
                f"main_occ={self.main_occ}, game_occ={self.game_occ}, side={self.side}, last_square={self.last_square})")
