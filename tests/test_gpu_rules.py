"""CUDA rules kernels (through the C ABI) vs the oracle rules engine — bit-exact.
PARITY UNPINNED w.r.t. utac itself: the oracle is our restatement (oracle/utac_rules.h)."""
import concurrent.futures as cf

import numpy as np
import pytest

import tacult_b200 as tb
from utac_gym.core import GameState, playout_many

pytestmark = pytest.mark.gpu


def random_positions(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        g = GameState()
        for _ in range(rng.randint(0, 75)):
            if g.is_terminal():
                break
            legal = np.flatnonzero(g.get_valid_mask())
            g.make_move(int(rng.choice(legal)))
            if g.current_player() != 1:
                gs = g._get_gs()
                gs.board = (~gs.board) & gs.occ
                gs.main_board = (~gs.main_board) & gs.main_occ
                gs.side = 1
                g = GameState(gs)
        out.append(g)
    return out


def oracle_status(g):
    if not g.is_terminal():
        return 0
    tv = g.terminal_value()
    return 2 if tv == 0 else (3 if tv == 1 else 1)


def test_playouts_match_oracle_1m():
    """>= 10^6 random playouts (~59M plies): checksum of every successor position, legal mask and status."""
    n, chunk, seed = 1_048_576, 65_536, 424242
    dev = tb.rules_playouts(seed, 0, n)
    d_sum = np.array([r.checksum for r in dev], dtype=np.uint64)
    d_plies = np.array([r.plies for r in dev])
    d_out = np.array([r.outcome for r in dev])

    def host(first):
        res = playout_many(seed, first, chunk)
        return (np.array([r.checksum for r in res], dtype=np.uint64), np.array([r.plies for r in res]),
                np.array([r.outcome for r in res]))
    with cf.ThreadPoolExecutor(max_workers=16) as ex:
        parts = list(ex.map(host, range(0, n, chunk)))
    assert (d_sum == np.concatenate([p[0] for p in parts])).all()
    assert (d_plies == np.concatenate([p[1] for p in parts])).all()
    assert (d_out == np.concatenate([p[2] for p in parts])).all()
    assert d_plies.sum() > 55_000_000


def test_playout_golden(golden):
    ref = golden("rules_playouts.npz")
    dev = tb.rules_playouts(int(ref["seed"]), int(ref["first_game"]), len(ref["checksum"]))
    assert (np.array([r.checksum for r in dev], dtype=np.uint64) == ref["checksum"]).all()
    assert (np.array([list(r.final_state) for r in dev], dtype=np.uint32) == ref["final_state"]).all()


def test_mask_status_obs_step_on_random_positions():
    pos = random_positions(3000, 5)
    states = np.stack([p.packed() for p in pos])
    mask = tb.rules_legal_mask(states)
    status = tb.rules_status(states)
    obs = tb.rules_encode_obs(states)
    for i, p in enumerate(pos):
        assert (mask[i] == p.get_valid_mask().astype(np.uint8)).all()
        assert status[i] == oracle_status(p)
        assert (obs[i].reshape(-1) == np.asarray(p.get_obs())).all()
    # one move from every live position
    rng = np.random.RandomState(9)
    live = [i for i, p in enumerate(pos) if not p.is_terminal()]
    acts = np.array([rng.choice(np.flatnonzero(mask[i])) for i in live], dtype=np.int8)
    nxt, nst = tb.rules_step(states[live], acts)
    for k, i in enumerate(live):
        g = GameState(pos[i])
        g.make_move(int(acts[k]))
        gs = g._get_gs()
        gs.board = (~gs.board) & gs.occ
        gs.main_board = (~gs.main_board) & gs.main_occ
        gs.side = 1
        g = GameState(gs)
        assert (nxt[k] == g.packed()).all() and nst[k] == oracle_status(g)


def test_edge_cases():
    assert tb.rules_legal_mask(np.zeros((0, 8), dtype=np.uint32)).shape == (0, 81)
    start = np.zeros((1, 8), dtype=np.uint32)
    assert tb.rules_legal_mask(start).sum() == 81
    with pytest.raises(tb._lib.TacultError) as err:
        nxt, _ = tb.rules_step(start, np.array([40], dtype=np.int8))
        tb.rules_step(nxt, np.array([0], dtype=np.int8))      # not in the forced sub-board
    assert err.value.code == tb._lib.TC_EILLEGAL
    # fields form and packed form agree
    g = random_positions(50, 11)
    gs = [p._get_gs() for p in g]
    packed = tb.pack_fields([x.board for x in gs], [x.occ for x in gs], [x.last_square for x in gs])
    assert (packed == np.stack([p.packed() for p in g])).all()
