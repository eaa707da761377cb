"""The product's host-side GameState (tacult_b200/gamestate.py over tc_gs_*, i.e. csrc/rules.cuh compiled for the host)
against the oracle rules engine, and the vectorised host glue of the drop-in (no GPU needed)."""
import sys

import numpy as np

from game_oracle import OracleGame
from tacult_b200 import _lib
from tacult_b200.engine import pack_gamestates
from tacult_b200.gamestate import GAMESTATE, GameState, install_as_utac_gym, pack_states, self_check_against
from tacult_b200.mcts import action_probs_from_counts
from utac_gym.core import GameState as OracleState
from utac_gym.core import playout_many


def mix64(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def product_playout(seed, game, max_plies=81):
    """oracle/utac_rules.c:utac_playout's specification, driven through the PRODUCT GameState."""
    m = (1 << 64) - 1
    s = GameState()
    total = mix64(seed ^ game)
    ply, first_to_move, outcome = 0, 1, 0
    while ply < max_plies and not s.is_terminal():
        legal = np.flatnonzero(s.get_valid_mask())
        r = mix64(seed ^ ((game * 0x9E3779B97F4A7C15) & m) ^ ((ply * 0xD1B54A32D192ED03) & m))
        s.make_move(int(legal[r % len(legal)]))
        if s.current_player() != 1:              # getCanonicalForm's colour flip (utac_game.py:97-107)
            gs = s._get_gs()
            gs.board = (~gs.board) & gs.occ
            gs.main_board = (~gs.main_board) & gs.main_occ
            gs.side = 1
            s = GameState(gs)
        first_to_move = -first_to_move
        ply += 1
        for w in s.packed():
            total = mix64(total ^ int(w))
        mask = s.get_valid_mask()
        words = [0, 0, 0]
        for a in np.flatnonzero(mask):
            words[int(a) >> 5] |= 1 << (int(a) & 31)
        for w in words:
            total = mix64(total ^ w)
        status = 0
        if s.is_terminal():
            tv = s.terminal_value()
            status = 2 if tv == 0 else (3 if tv == 1 else 1)
            if tv != 0:
                outcome = first_to_move if tv == 1 else -first_to_move
        total = mix64(total ^ status)
    return total, ply, outcome, s.packed()


def test_playout_checksums_match_oracle_rules():
    """Every successor position, legal mask and terminal status of 300 seeded games, folded into the checksum the
    oracle's C playout driver (and the CUDA one) computes."""
    want = playout_many(11, 0, 300)
    for g in range(300):
        total, plies, outcome, final = product_playout(11, g)
        assert (total, plies, outcome) == (want[g].checksum, want[g].plies, want[g].outcome), g
        assert list(final) == list(want[g].final_state)


def test_every_method_matches_oracle_ply_by_ply():
    assert self_check_against(OracleState, games=120, seed=3, device_hooks=False) > 1000


def test_value_semantics_and_gamestate_round_trip():
    a = GameState()
    a.make_move(40)
    b = GameState(a)
    b.make_move(int(np.flatnonzero(b.get_valid_mask())[0]))
    assert a._get_gs().last_square == 4 and a.current_player() == -1 and b.current_player() == 1
    assert int(a.get_valid_mask().sum()) == 8           # sub-board 4 minus the centre
    gs = b._get_gs()
    assert isinstance(gs, GAMESTATE) and gs.board.dtype == np.uint16
    c = GameState(gs)
    assert bytes(c._c) == bytes(b._c)
    try:
        a.make_move(40)
        raise AssertionError("occupied cell accepted")
    except ValueError:
        pass
    assert GameState.from_packed(GameState().packed()).packed().tolist() == [0] * 8


def test_pack_states_one_call_matches_field_loop():
    game = OracleGame()
    rng = np.random.RandomState(5)
    mine, theirs = [], []
    for i in range(300):
        a, b = GameState(), OracleState()
        for _ in range(rng.randint(0, 60)):
            if a.is_terminal():
                break
            mv = int(rng.choice(np.flatnonzero(a.get_valid_mask())))
            a.make_move(mv)
            b.make_move(mv)
        if a.current_player() != 1 and not a.is_terminal():
            b = game.getCanonicalForm(b, -1)
            a = GameState(b._get_gs())
        mine.append(a if i % 17 else None)
        theirs.append(b if i % 17 else None)
    fast = pack_states(mine)
    assert fast is not None
    slow = pack_gamestates(theirs)                      # oracle objects: the per-object `_get_gs()` path
    assert np.array_equal(fast[1], slow[1])
    on = fast[1].astype(bool)
    assert np.array_equal(fast[0][on], slow[0][on])
    assert pack_states([OracleState()]) is None         # foreign objects are left to the field loop
    # a live board that is not in canonical form is an error, as in the field loop
    bad = GameState()
    bad.make_move(0)
    try:
        pack_states([bad])
        raise AssertionError("non-canonical live board accepted")
    except _lib.TacultError as exc:
        assert exc.code == _lib.TC_EINVAL


def test_install_as_utac_gym_serves_the_reference_imports():
    saved = {k: sys.modules.get(k) for k in ("utac_gym", "utac_gym.core", "utac_gym.core.types")}
    try:
        assert install_as_utac_gym(force=True)
        from utac_gym.core import GameState as G
        from utac_gym.core.types import GAMESTATE as S
        assert G is GameState and S is GAMESTATE
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def reference_tail(counts_all, temps):
    """getActionProbs' per-env tail as the reference writes it (/root/reference/src/tacult/base/mcts.py:200-214)."""
    n_env, n_act = counts_all.shape
    probs = np.zeros((n_env, n_act))
    for i in range(n_env):
        counts, temp = counts_all[i], temps[i]
        if temp == 0:
            best = np.array(np.argwhere(counts == np.max(counts))).ravel()
            probs[i][np.random.choice(best)] = 1
        else:
            powered = counts ** (1. / temp)
            with np.errstate(invalid="ignore", divide="ignore"):
                probs[i] = powered / float(np.sum(powered))
    return probs


def test_vectorised_action_probs_equal_the_per_env_loop_draw_for_draw():
    rng = np.random.RandomState(1)
    for case in range(4):
        counts = rng.randint(0, [3, 5, 200, 50][case], size=(257, 81)).astype(np.int64)
        counts[::9] = 0                                  # finished envs: NaN rows / random one-hot
        temps = [np.zeros(257), np.ones(257), rng.randint(0, 2, 257).astype(bool),
                 rng.choice([0.0, 1.0, 0.5], 257)][case]
        np.random.seed(100 + case)
        want = reference_tail(counts, temps)
        state_want = np.random.get_state()
        np.random.seed(100 + case)
        got = action_probs_from_counts(counts, temps)
        state_got = np.random.get_state()
        assert np.array_equal(want, got, equal_nan=True)
        assert np.array_equal(state_want[1], state_got[1]) and state_want[2] == state_got[2]


def test_choose_actions_equals_the_numpy_formulation():
    """tc_choose_actions (the actor's move choice, coach.py:125-140) against cumsum / arg-max in numpy, draw for draw,
    including zero rows (finished envs), single-action rows and exact ties."""
    from tacult_b200.engine import choose_actions
    rng = np.random.RandomState(3)
    counts = rng.randint(0, 40, size=(3000, 81)).astype(np.uint32)
    counts[rng.random_sample(counts.shape) < 0.6] = 0
    counts[::11] = 0
    counts[5::11, :] = 0
    counts[5::11, 17] = 200
    counts[7::13, :] = 3                                    # ties everywhere: the lowest index wins at tau = 0
    tau = rng.randint(0, 2, 3000).astype(bool)
    u = rng.random_sample(3000)
    u[:50] = np.nextafter(1.0, 0.0)                          # the largest draw still picks a visited action
    want = counts.argmax(1)
    cum = np.cumsum(counts, axis=1, dtype=np.int64)
    draw = (u * cum[:, -1]).astype(np.int64)
    want = np.where(tau, (cum > draw[:, None]).argmax(1), want)
    got = choose_actions(counts, tau, u)
    assert got.dtype == np.int8 and np.array_equal(got, want)
    sampled = tau & (cum[:, -1] > 0)
    assert (counts[np.flatnonzero(sampled), got[sampled]] > 0).all()
    assert np.array_equal(choose_actions(counts, np.zeros(3000, bool)), counts.argmax(1))      # u not needed at tau = 0
