"""Oracle rules engine: hand-derived known answers, frozen playout checksums, and the product's
bit-parallel rules (tacult_b200/csrc/rules.cuh compiled for the host) against the scalar oracle.
The rules engine is PARITY UNPINNED with respect to utac (oracle/utac_rules.h)."""
import os
import subprocess

import numpy as np

from conftest import ROOT
from game_oracle import OracleGame
from utac_gym.core import GameState, playout_many


def a_of(r, c):
    return 9 * r + c


def test_initial_position():
    g = GameState()
    assert g.current_player() == 1 and not g.is_terminal()
    assert g.get_valid_mask().sum() == 81
    gs = g._get_gs()
    assert gs.last_square == -1 and gs.side == 1 and int(np.sum(gs.occ)) == 0
    obs = np.asarray(g.get_obs()).reshape(4, 9, 9)
    assert obs[0].sum() == 0 and obs[1].sum() == 0 and obs[2].sum() == 81 and obs[3].sum() == 81


def test_bit_layout_and_forced_subboard():
    # move at global (4,4): sub-board 4, cell 4 -> opponent is sent to sub-board 4 (8 empty cells left)
    g = GameState()
    g.make_move(a_of(4, 4))
    gs = g._get_gs()
    assert gs.occ[4] == 1 << 4 and gs.board[4] == 1 << 4 and gs.last_square == 4 and gs.side == -1
    mask = g.get_valid_mask().reshape(9, 9)
    assert mask.sum() == 8 and mask[3:6, 3:6].sum() == 8 and mask[4, 4] == 0
    # move at global (3,5): sub 4, cell (0,2)=2 -> next forced sub-board 2 (rows 0-2, cols 6-8)
    g.make_move(a_of(3, 5))
    gs = g._get_gs()
    assert gs.occ[4] == (1 << 4) | (1 << 2) and gs.board[4] == 1 << 4 and gs.last_square == 2 and gs.side == 1
    mask = g.get_valid_mask().reshape(9, 9)
    assert mask.sum() == 9 and mask[0:3, 6:9].sum() == 9


def _play(g, moves):
    for r, c in moves:
        g.make_move(a_of(r, c))


def test_subboard_win_closes_it_and_free_move():
    g = GameState()
    # O collects cells 1,3,6,4,2 of sub-board 0 (anti-diagonal 2,4,6) while X is bounced around
    _play(g, [(0, 0), (0, 1)])          # X sub0 cell0 -> O forced to sub0, plays cell1 -> X forced to sub1
    assert g._get_gs().last_square == 1
    _play(g, [(0, 3), (1, 0)])          # X sub1 cell0 -> O to sub0 cell3 -> X to sub3
    _play(g, [(3, 0), (2, 0)])          # X sub3 cell0 -> O sub0 cell6 -> X to sub6
    _play(g, [(6, 0), (1, 1)])          # X sub6 cell0 -> O sub0 cell4 -> X to sub4
    assert g._get_gs().main_occ == 0
    _play(g, [(3, 3), (0, 2)])          # X sub4 cell0 -> O sub0 cell2: line 2,4,6
    gs = g._get_gs()
    assert gs.main_occ == 1 and gs.main_board == 0 and gs.game_occ == 1      # O won sub-board 0
    assert gs.last_square == 2
    assert g.get_valid_mask().reshape(9, 9)[0:3, 6:9].sum() == 9             # X is sent to sub 2 (open)
    _play(g, [(0, 6)])                  # X sub2 cell0 -> O sent to closed sub 0 => free move anywhere open
    mask = g.get_valid_mask().reshape(9, 9)
    assert mask[0:3, 0:3].sum() == 0    # closed sub-board accepts nothing
    stones_elsewhere = int(sum(bin(int(x)).count("1") for x in g._get_gs().occ[1:]))
    assert mask.sum() == 72 - stones_elsewhere


def test_canonical_flip_and_key():
    game = OracleGame()
    g = GameState()
    g.make_move(40)
    canon = game.getCanonicalForm(g, -1)
    gs = canon._get_gs()
    assert gs.side == 1 and gs.board[4] == 0 and gs.occ[4] == 1 << 4
    key = game.stringRepresentation(canon)
    assert len(key) == 81 + 81 + 2 and key.endswith("14")
    assert game.stringRepresentation(GameState()).endswith("1-1") and len(game.stringRepresentation(GameState())) == 165
    assert game.getGameEnded(canon, 1) == 0
    assert (game.getValidMoves(g, 1) == 0).all() and game.getValidMoves(canon, 1).sum() == 8


def test_full_game_terminal_values():
    res = playout_many(99, 0, 3000)
    outcomes = np.array([r.outcome for r in res])
    plies = np.array([r.plies for r in res])
    assert set(np.unique(outcomes)) == {-1, 0, 1}          # draws exist (utac_game.py:78-79)
    assert plies.min() >= 17 and plies.max() <= 81
    for r in res[:200]:                                    # a finished position has no legal moves
        g = GameState.from_packed(np.array(list(r.final_state), dtype=np.uint32))
        assert g.is_terminal() and g.get_valid_mask().sum() == 0
        tv = g.terminal_value()
        assert (tv == 0) == (r.outcome == 0)
        assert tv in (0, -1)     # the side to move can never be the winner


def test_playout_checksums_frozen(golden):
    ref = golden("rules_playouts.npz")
    res = playout_many(int(ref["seed"]), int(ref["first_game"]), len(ref["checksum"]))
    assert (np.array([r.checksum for r in res], dtype=np.uint64) == ref["checksum"]).all()
    assert (np.array([r.plies for r in res]) == ref["plies"]).all()
    assert (np.array([r.outcome for r in res]) == ref["outcome"]).all()
    assert (np.array([list(r.final_state) for r in res], dtype=np.uint32) == ref["final_state"]).all()


def test_symmetries_match_reference_golden(golden):
    ref = golden("symmetries.npz")
    game = OracleGame()
    for k in range(len(ref["states"])):
        pos = GameState.from_packed(ref["states"][k])
        syms = game.getSymmetries(pos, np.arange(81, dtype=np.float64))
        assert len(syms) == 8
        for j, (b, p) in enumerate(syms):
            assert (b.packed() == ref["images"][k][j]).all()
            assert (p.astype(np.int16) == ref["perms"][k][j]).all()


def test_product_bitboard_rules_match_oracle_on_host(tmp_path):
    exe = str(tmp_path / "host_rules_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe,
                           os.path.join(ROOT, "tests", "host_rules_check.cpp"),
                           os.path.join(ROOT, "oracle", "utac_rules.c")])
    out = subprocess.run([exe, "8000", "777"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
