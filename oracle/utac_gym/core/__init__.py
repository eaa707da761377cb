"""ORACLE — TEST INFRASTRUCTURE ONLY.

``GameState``: the Python surface of the external rules engine as inferred from
the reference's call sites:
  GameState() / GameState(other) / GameState(GAMESTATE)   utac_game.py:18,45,100,107,174,218
  make_move                                               utac_game.py:46
  get_valid_mask                                          utac_game.py:62
  is_terminal / terminal_value                            utac_game.py:75-81
  current_player                                          utac_game.py:244
  get_obs                                                 utac_game.py:248, base/utils.py:8
  _get_gs                                                 utac_game.py:100,174,240
  print                                                   base/coach.py:68,234
Backed by the plain-C restatement in oracle/utac_rules.c through ctypes.
"""
import ctypes
import os
import subprocess

import numpy as np

from .types import GAMESTATE

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.normpath(os.path.join(_HERE, "..", ".."))


class _CState(ctypes.Structure):
    _fields_ = [
        ("board", ctypes.c_uint16 * 9),
        ("occ", ctypes.c_uint16 * 9),
        ("main_board", ctypes.c_uint16),
        ("main_occ", ctypes.c_uint16),
        ("game_occ", ctypes.c_uint16),
        ("side", ctypes.c_int8),
        ("last_square", ctypes.c_int8),
    ]


class PlayoutResult(ctypes.Structure):
    _fields_ = [
        ("checksum", ctypes.c_uint64),
        ("final_state", ctypes.c_uint32 * 8),
        ("plies", ctypes.c_int32),
        ("outcome", ctypes.c_int32),
    ]


def _load():
    so = os.path.join(_ORACLE_DIR, "liboracle_utac.so")
    src = os.path.join(_ORACLE_DIR, "utac_rules.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "liboracle_utac.so"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    P = ctypes.POINTER(_CState)
    lib.utac_init.argtypes = [P]
    lib.utac_rederive.argtypes = [P]
    lib.utac_make_move.argtypes = [P, ctypes.c_int]
    lib.utac_make_move.restype = ctypes.c_int
    lib.utac_valid_mask.argtypes = [P, ctypes.c_void_p]
    lib.utac_valid_mask.restype = ctypes.c_int
    lib.utac_is_terminal.argtypes = [P]
    lib.utac_is_terminal.restype = ctypes.c_int
    lib.utac_terminal_value.argtypes = [P]
    lib.utac_terminal_value.restype = ctypes.c_int
    lib.utac_obs.argtypes = [P, ctypes.c_void_p]
    lib.utac_flip.argtypes = [P]
    lib.utac_pack_canonical.argtypes = [P, ctypes.c_void_p]
    lib.utac_unpack_canonical.argtypes = [ctypes.c_void_p, P]
    lib.utac_mix64.argtypes = [ctypes.c_uint64]
    lib.utac_mix64.restype = ctypes.c_uint64
    lib.utac_playout_many.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(PlayoutResult)]
    return lib


LIB = _load()


class GameState:
    __slots__ = ("_c",)

    def __init__(self, other=None):
        self._c = _CState()
        if other is None:
            LIB.utac_init(ctypes.byref(self._c))
        elif isinstance(other, GameState):
            ctypes.memmove(ctypes.byref(self._c), ctypes.byref(other._c), ctypes.sizeof(_CState))
        else:  # GAMESTATE-like: fields are taken as given
            c = self._c
            for i in range(9):
                c.board[i] = int(other.board[i]) & 0x1FF
                c.occ[i] = int(other.occ[i]) & 0x1FF
            c.main_board = int(other.main_board) & 0x1FF
            c.main_occ = int(other.main_occ) & 0x1FF
            c.game_occ = int(other.game_occ) & 0x1FF
            c.side = 1 if int(other.side) == 1 else -1
            c.last_square = int(other.last_square)

    # --- rules -----------------------------------------------------------
    def make_move(self, action):
        if LIB.utac_make_move(ctypes.byref(self._c), int(action)) != 0:
            raise ValueError(f"illegal move {action}")

    def get_valid_mask(self):
        buf = np.empty(81, dtype=np.uint8)
        LIB.utac_valid_mask(ctypes.byref(self._c), buf.ctypes.data)
        return buf.astype(np.float64)

    def is_terminal(self):
        return bool(LIB.utac_is_terminal(ctypes.byref(self._c)))

    def terminal_value(self):
        return int(LIB.utac_terminal_value(ctypes.byref(self._c)))

    def current_player(self):
        return int(self._c.side)

    def get_obs(self):
        buf = np.empty(324, dtype=np.float32)
        LIB.utac_obs(ctypes.byref(self._c), buf.ctypes.data)
        return buf

    def _get_gs(self):
        c = self._c
        return GAMESTATE(board=list(c.board), occ=list(c.occ), main_board=c.main_board, main_occ=c.main_occ,
                         game_occ=c.game_occ, side=c.side, last_square=c.last_square)

    # --- helpers used only by tests --------------------------------------
    def packed(self):
        """8 x u32 canonical device packing (include/tacult_b200.h)."""
        out = np.zeros(8, dtype=np.uint32)
        LIB.utac_pack_canonical(ctypes.byref(self._c), out.ctypes.data)
        return out

    @classmethod
    def from_packed(cls, words):
        g = cls()
        w = np.ascontiguousarray(words, dtype=np.uint32)
        LIB.utac_unpack_canonical(w.ctypes.data, ctypes.byref(g._c))
        return g

    def print(self):
        c = self._c
        rows = []
        for r in range(9):
            line = []
            for col in range(9):
                sub, cell = (r // 3) * 3 + col // 3, (r % 3) * 3 + col % 3
                if (c.board[sub] >> cell) & 1:
                    line.append("X")
                elif (c.occ[sub] >> cell) & 1:
                    line.append("O")
                else:
                    line.append(".")
                if col % 3 == 2 and col != 8:
                    line.append("|")
            rows.append(" ".join(line))
            if r % 3 == 2 and r != 8:
                rows.append("-" * 21)
        print("\n".join(rows))
        print(f"side={c.side} last_square={c.last_square}")


def playout_many(seed, first_game, n, max_plies=81):
    res = (PlayoutResult * n)()
    LIB.utac_playout_many(seed, first_game, n, max_plies, res)
    return res
