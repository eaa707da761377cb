"""Host-side `GameState` / `GAMESTATE`: the two names the reference imports from the external `utac_gym`
package (`from utac_gym.core import GameState`, `from utac_gym.core.types import GAMESTATE`:
/root/reference/src/tacult/utac_game.py:2-3, base/coach.py:17-18), with exactly the surface its call sites
use:

    GameState() / GameState(other) / GameState(GAMESTATE)    utac_game.py:18,45,100,107,174,218
    .make_move(action)                                       utac_game.py:46
    .get_valid_mask()                                        utac_game.py:62
    .is_terminal() / .terminal_value()                       utac_game.py:75-81
    .current_player()                                        utac_game.py:244
    .get_obs()                                               utac_game.py:248, base/utils.py:8
    ._get_gs()                                               utac_game.py:100,174,240
    .print()                                                 base/coach.py:68,234

The rules are the library's own bit-parallel implementation (csrc/rules.cuh, compiled for the host behind
`tc_gs_*`), the very code the CUDA search and the device rules hooks run — so a `Coach` driven by these
objects and the GPU engine can never disagree about a position.  `install_as_utac_gym()` registers the
classes under the `utac_gym` module names when the real package is absent, so the reference's own files
import unchanged.

Rule choices the reference does not pin (closed sub-boards, free move, draw criterion, plane 2/3 content)
follow SURVEY.md §10.3; `self_check_against(other_module)` compares them with a live `utac_gym` when one is
importable and fails loudly on the first difference.
"""
import ctypes
import sys
import types

import numpy as np

from . import _lib
from ._lib import GameStateStruct

_SIZE = ctypes.sizeof(GameStateStruct)


class GAMESTATE:
    """Raw struct view the reference reads and mutates (utac_game.py:100-106,174-216,240-241)."""
    __slots__ = ("board", "occ", "main_board", "main_occ", "game_occ", "side", "last_square")

    def __init__(self, board=None, occ=None, main_board=0, main_occ=0, game_occ=0, side=1, last_square=-1):
        self.board = np.zeros(9, dtype=np.uint16) if board is None else np.array(board, dtype=np.uint16)
        self.occ = np.zeros(9, dtype=np.uint16) if occ is None else np.array(occ, dtype=np.uint16)
        self.main_board, self.main_occ, self.game_occ = int(main_board), int(main_occ), int(game_occ)
        self.side, self.last_square = int(side), int(last_square)

    def __repr__(self):
        return (f"GAMESTATE(board={self.board.tolist()}, occ={self.occ.tolist()}, main_board={self.main_board}, "
                f"main_occ={self.main_occ}, game_occ={self.game_occ}, side={self.side}, "
                f"last_square={self.last_square})")


class GameState:
    __slots__ = ("_c",)
    _lib = None

    def __init__(self, other=None):
        lib = GameState._lib
        if lib is None:
            lib = GameState._lib = _lib.load()
        c = self._c = GameStateStruct()
        if other is None:
            lib.tc_gs_init(ctypes.byref(c))
        elif isinstance(other, GameState):
            ctypes.memmove(ctypes.byref(c), ctypes.byref(other._c), _SIZE)
        else:
            # GAMESTATE-like: the key fields are taken as given, the derived ones recomputed (utac_game.py:222-229
            # shows board, occ, side and last_square determine the position)
            for i in range(9):
                c.board[i] = int(other.board[i]) & 0x1FF
                c.occ[i] = int(other.occ[i]) & 0x1FF
            c.side = 1 if int(other.side) == 1 else -1
            c.last_square = int(other.last_square)
            lib.tc_gs_rederive(ctypes.byref(c))

    def make_move(self, action):
        if self._lib.tc_gs_make_move(ctypes.byref(self._c), int(action)) != _lib.TC_OK:
            raise ValueError(f"illegal move {action}")

    def get_valid_mask(self):
        out = np.empty(81, dtype=np.float64)
        self._lib.tc_gs_valid_mask(ctypes.byref(self._c), out.ctypes.data)
        return out

    def is_terminal(self):
        return bool(self._lib.tc_gs_is_terminal(ctypes.byref(self._c)))

    def terminal_value(self):
        return int(self._lib.tc_gs_terminal_value(ctypes.byref(self._c)))

    def current_player(self):
        return int(self._c.side)

    def get_obs(self):
        out = np.empty(324, dtype=np.float32)
        self._lib.tc_gs_obs(ctypes.byref(self._c), out.ctypes.data)
        return out

    def _get_gs(self):
        c = self._c
        return GAMESTATE(list(c.board), list(c.occ), c.main_board, c.main_occ, c.game_occ, c.side, c.last_square)

    def packed(self):
        """Canonical packing (include/tacult_b200.h); only meaningful when side == 1."""
        out = np.zeros(8, dtype=np.uint32)
        active = np.ones(1, dtype=np.uint8)
        _lib.check(self._lib.tc_gs_pack(1, ctypes.addressof(self._c), out.ctypes.data, active.ctypes.data))
        return out

    @classmethod
    def from_packed(cls, words):
        w = [int(x) for x in np.asarray(words, dtype=np.uint32).reshape(8)]
        gs = GAMESTATE(side=1, last_square=(w[6] & 0xF) - 1)
        for i in range(9):
            gs.board[i] = (w[i // 3] >> (9 * (i % 3))) & 0x1FF
            gs.occ[i] = (w[3 + i // 3] >> (9 * (i % 3))) & 0x1FF
        return cls(gs)

    def print(self):
        c = self._c
        lines = []
        for r in range(9):
            cells = []
            for col in range(9):
                sub, cell = (r // 3) * 3 + col // 3, (r % 3) * 3 + col % 3
                stone = "X" if (c.board[sub] >> cell) & 1 else ("O" if (c.occ[sub] >> cell) & 1 else ".")
                cells.append(stone + (" |" if col in (2, 5) else ""))
            lines.append(" ".join(cells))
            if r in (2, 5):
                lines.append("-" * 21)
        print("\n".join(lines))
        print(f"side={c.side} last_square={c.last_square}")


def pack_states(boards):
    """list of GameState (None allowed) -> (packed u32[n,8], active u8[n]) through ONE C call (tc_gs_pack).
    Returns None when the list holds foreign objects (e.g. the real utac_gym's GameState)."""
    n = len(boards)
    active = np.ones(n, dtype=np.uint8)
    blank = None
    chunks = []
    for i, b in enumerate(boards):
        if b is None:
            active[i] = 0
            if blank is None:
                blank = GameStateStruct()
            chunks.append(blank)
        elif type(b) is GameState:
            chunks.append(b._c)
        else:
            return None
    raw = b"".join(chunks)
    out = np.empty((n, 8), dtype=np.uint32)
    _lib.check(_lib.load().tc_gs_pack(n, raw, out.ctypes.data, active.ctypes.data))
    return out, active


def install_as_utac_gym(force=False):
    """Registers `utac_gym`, `utac_gym.core` and `utac_gym.core.types` in sys.modules backed by this module, so
    that the reference's `from utac_gym.core import GameState` resolves here.  Leaves a real utac_gym alone
    unless `force`."""
    if not force:
        try:
            import utac_gym.core  # noqa: F401
            return False
        except ImportError:
            pass
    pkg, core, tps = types.ModuleType("utac_gym"), types.ModuleType("utac_gym.core"), types.ModuleType("utac_gym.core.types")
    pkg.__path__, core.__path__ = [], []
    core.GameState, tps.GAMESTATE = GameState, GAMESTATE
    pkg.core, core.types = core, tps
    pkg._tacult_b200_shim = True
    sys.modules["utac_gym"], sys.modules["utac_gym.core"], sys.modules["utac_gym.core.types"] = pkg, core, tps
    return True


def self_check_against(foreign_game_state, games=64, seed=0, device_hooks=True):
    """Plays `games` random games with `foreign_game_state` (e.g. the real `utac_gym.core.GameState`) and this
    module's GameState side by side and compares, at every ply, legal mask, terminal flag / value, current player,
    observation planes and raw fields; with `device_hooks` the CUDA hooks (tc_rules_legal_mask / encode_obs / step)
    are compared on the canonical positions too.  Raises AssertionError on the first difference: the rule choices
    and plane semantics the reference does not pin are then NOT those of that rules engine, and
    reference-trained checkpoints must not be used with this engine."""
    from .engine import rules_encode_obs, rules_legal_mask
    rng = np.random.RandomState(seed)
    canon = []
    for g in range(games):
        a, b = foreign_game_state(), GameState()
        for ply in range(82):
            ga, gb = a._get_gs(), b._get_gs()
            where = f"game {g} ply {ply}"
            assert list(np.asarray(ga.board)) == list(gb.board) and list(np.asarray(ga.occ)) == list(gb.occ), where
            assert int(ga.last_square) == gb.last_square and (int(ga.side) == 1) == (gb.side == 1), where
            assert int(ga.main_board) == gb.main_board and int(ga.main_occ) == gb.main_occ, f"{where}: main_*"
            assert int(ga.game_occ) == gb.game_occ, f"{where}: game_occ"
            ma, mb = np.asarray(a.get_valid_mask(), dtype=np.float64), b.get_valid_mask()
            assert np.array_equal(ma != 0, mb != 0), f"{where}: legal masks differ"
            assert bool(a.is_terminal()) == b.is_terminal(), f"{where}: is_terminal"
            assert np.array_equal(np.asarray(a.get_obs(), dtype=np.float32).ravel(), b.get_obs()), f"{where}: planes"
            if b.is_terminal():
                assert int(a.terminal_value()) == b.terminal_value(), f"{where}: terminal_value"
                break
            if b.current_player() == 1:
                canon.append(b.packed())
            move = int(rng.choice(np.flatnonzero(mb)))
            a.make_move(move)
            b.make_move(move)
    if device_hooks and canon:
        states = np.stack(canon)
        host = [GameState.from_packed(s) for s in states]
        assert np.array_equal(rules_legal_mask(states) != 0, np.stack([h.get_valid_mask() for h in host]) != 0)
        assert np.array_equal(rules_encode_obs(states).reshape(len(host), 324), np.stack([h.get_obs() for h in host]))
    return len(canon)
