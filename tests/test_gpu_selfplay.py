"""Device-resident self-play driver vs Coach.executeEpisode semantics
(/root/reference/src/tacult/base/coach.py:96-157), checked with the oracle rules and search."""
import logging

import numpy as np
import pytest

import tacult_b200 as tb
from evaluators import HashEvaluator
from game_oracle import OracleGame
from mcts_oracle import OracleVectorizedMCTS
from utac_gym.core import GameState

pytestmark = pytest.mark.gpu
logging.disable(logging.ERROR)


def run_games(envs, sims, threshold, seed, salt):
    eng = tb.Engine(envs, sims * 82 + 64, hash_salt=salt)
    eng.selfplay_begin(sims, 2.4, threshold, seed=seed, auto_reset=False)
    for _ in range(81):
        eng.selfplay_step(1)
        if eng.selfplay_stats()["games_finished"] == envs:
            break
    eng.selfplay_sync()
    return eng, eng.selfplay_drain()


def test_trajectories_follow_the_rules_and_the_search():
    envs, sims, threshold, salt = 6, 24, 6, 31
    eng, recs = run_games(envs, sims, threshold, seed=5, salt=salt)
    stats = eng.selfplay_stats()
    assert stats["games_finished"] == envs and stats["examples"] == 0 and len(recs) == stats["plies"]
    assert stats["sims"] == stats["plies"] * sims
    assert stats["first_player_wins"] + stats["second_player_wins"] + stats["draws"] == envs
    game = OracleGame()
    args = type("A", (dict,), {"__getattr__": dict.__getitem__})(numEnvs=envs, numMCTSSims=sims, cpuct=2.4)
    cpu = OracleVectorizedMCTS(game, HashEvaluator(salt), args)
    by_env = {e: sorted([r for r in recs if r["env"] == e], key=lambda r: r["ply"]) for e in range(envs)}
    max_len = max(len(v) for v in by_env.values())
    boards = [GameState() for _ in range(envs)]
    players = [1] * envs
    for ply in range(max_len):
        canon = [game.getCanonicalForm(boards[e], players[e]) if ply < len(by_env[e]) else None for e in range(envs)]
        for _ in range(sims):
            cpu.search(canon)
        for e in range(envs):
            if ply >= len(by_env[e]):
                continue
            r = by_env[e][ply]
            assert (r["state"] == canon[e].packed()).all() and r["player"] == players[e] and r["ply"] == ply
            want = cpu.root_counts([c if c is not None else GameState() for c in canon])[e]
            assert (r["counts"] == want).all()                       # same visit counts as the CPU search
            assert r["tau_one"] == (1 if ply + 1 < threshold else 0)  # coach.py:124-125
            a = int(r["action"])
            assert want[a] > 0 and (r["tau_one"] or want[a] == want.max())
            boards[e], players[e] = game.getNextState(boards[e], players[e], a)
    for e in range(envs):
        result = game.getGameEnded(boards[e], players[e])             # coach.py:143
        assert result != 0
        for r in by_env[e]:
            z = result * ((-1) ** (int(r["player"]) != players[e]))   # coach.py:149
            got = float(r["z_sign"]) * (1e-4 if r["finished"] == 2 else 1.0)
            assert got == pytest.approx(z) and r["finished"] in (1, 2)


def test_auto_reset_keeps_all_envs_busy():
    envs, sims = 64, 8
    eng = tb.Engine(envs, sims * 82 + 64, hash_salt=2)
    eng.selfplay_begin(sims, 2.4, 10, seed=1, auto_reset=True)
    recs = []
    for _ in range(3):
        eng.selfplay_step(50)
        eng.selfplay_sync()
        recs.append(eng.selfplay_drain())
    recs = np.concatenate(recs)
    st = eng.selfplay_stats()
    assert st["plies"] == envs * 150 and st["games_finished"] >= envs
    assert len(recs) > 0 and (recs["finished"] > 0).all()
    assert set(np.unique(recs["game"])) >= {0, 1}
    # every finished game is complete: plies 0..n-1 of (env, game) are all there
    for env in range(0, envs, 7):
        mine = recs[(recs["env"] == env) & (recs["game"] == 0)]
        assert sorted(mine["ply"]) == list(range(len(mine))) and len(mine) >= 17


def test_record_buffer_overflow_is_loud_and_discard_mode_is_silent():
    envs, sims = 32, 2
    eng = tb.Engine(envs, sims * 82 + 64, hash_salt=2)
    eng.selfplay_begin(sims, 2.4, 10, seed=1, auto_reset=True)
    eng.selfplay_step(600)                  # ~19k records into a buffer of 32*256
    eng.selfplay_sync()
    with pytest.raises(tb._lib.TacultError) as err:
        eng.selfplay_drain()
    assert err.value.code == tb._lib.TC_ECAPACITY
    assert len(eng.selfplay_drain()) == 0   # buffer restarted empty
    eng.selfplay_configure(sims, 2.4, 10, seed=1, auto_reset=True, discard_records=True)
    eng.selfplay_step(600)
    eng.selfplay_sync()
    assert len(eng.selfplay_drain()) == 0


def test_records_to_training_examples_match_coach_format():
    """(obs, pi, z) tuples incl. the 8 symmetries, as coach.py:136-149 builds them, from the device records."""
    from tacult_b200.selfplay import arrays_to_examples, records_to_arrays
    envs, sims, threshold, salt = 4, 16, 5, 77
    eng, recs = run_games(envs, sims, threshold, seed=9, salt=salt)
    obs, pi, z = records_to_arrays(recs, augment=True)
    assert obs.shape == (8 * len(recs), 4, 9, 9) and pi.shape == (8 * len(recs), 81) and z.shape == (8 * len(recs),)
    game = OracleGame()
    for k in range(0, len(recs), 5):
        r = recs[k]
        pos = GameState.from_packed(r["state"])
        counts = r["counts"].astype(np.float64)
        want_pi = counts / counts.sum() if r["tau_one"] else np.eye(81)[int(r["action"])]
        syms = game.getSymmetries(pos, want_pi)
        for j, (b, p) in enumerate(syms):
            assert np.array_equal(obs[8 * k + j], np.asarray(b.get_obs()).reshape(4, 9, 9))
            assert np.allclose(pi[8 * k + j], p.astype(np.float32))
        want_z = float(r["z_sign"]) * (1e-4 if r["finished"] == 2 else 1.0)
        assert np.allclose(z[8 * k:8 * k + 8], want_z)
    ex = arrays_to_examples(obs, pi, z)
    assert len(ex) == 8 * len(recs) and tuple(ex[0][0].shape) == (4, 9, 9) and tuple(ex[0][1].shape) == (81,)


def test_selfplay_driver_generates_numenvs_games():
    from tacult_b200.selfplay import SelfPlayDriver

    class A(dict):
        __getattr__ = dict.__getitem__

    class HashNet:
        tc_hash_params = (5, 0, 0)
    drv = SelfPlayDriver(HashNet(), A(numEnvs=16, numMCTSSims=8, cpuct=2.4, tempThreshold=10))
    recs, stats = drv.play_records(seed=3)
    assert stats["games_finished"] == 16 and len(recs) == stats["plies"]      # exactly numEnvs games (coach.py:175)
    assert sorted(set(recs["env"])) == list(range(16))


def test_device_drain_and_example_kernel_match_the_host_path():
    """tc_selfplay_drain_dev leaves the records in HBM and tc_examples_from_records_dev builds (obs, pi, z) with the
    eight symmetries there; both must equal the host path (tc_selfplay_drain + records_to_arrays, itself checked
    against getSymmetries above) bit for bit — for all examples in order and for an arbitrary index list."""
    import torch
    from tacult_b200.distributed import DeviceReplay, DeviceTrajectoryExchange
    from tacult_b200.selfplay import records_to_arrays
    envs, sims, threshold, salt = 24, 12, 6, 41
    _, host_recs = run_games(envs, sims, threshold, seed=13, salt=salt)
    eng = tb.Engine(envs, sims * 82 + 64, hash_salt=salt)
    eng.selfplay_begin(sims, 2.4, threshold, seed=13, auto_reset=False)
    eng.selfplay_step(81)
    eng.selfplay_sync()
    ex = DeviceTrajectoryExchange(eng, "cuda:0")
    recs_dev, counts = ex.gather()
    assert counts == [len(host_recs)] and recs_dev.is_cuda
    got = recs_dev.cpu().numpy().reshape(-1).view(tb._lib.EXAMPLE_DTYPE)
    # games that end on the same ply append their records in whatever order their warps reach the atomic counter: sort
    # both by (env, ply) and keep the device order for the example checks below
    key_got, key_host = np.lexsort((got["ply"], got["env"])), np.lexsort((host_recs["ply"], host_recs["env"]))
    for name in tb._lib.EXAMPLE_DTYPE.names:             # field by field: numpy does not carry padding bytes through copies
        assert np.array_equal(got[name][key_got], host_recs[name][key_host]), name
    host_recs = got.copy()
    assert eng.selfplay_drain_dev(ex.send.data_ptr(), ex.capacity) == 0          # drained
    replay = DeviceReplay(eng, recs_dev)
    obs, pi, z = (t.cpu().numpy() for t in replay.examples())
    want_obs, want_pi, want_z = records_to_arrays(host_recs, augment=True)
    assert np.array_equal(obs, want_obs) and np.array_equal(pi, want_pi) and np.array_equal(z, want_z)
    idx = torch.from_numpy(np.random.RandomState(0).randint(0, replay.n_examples, size=1000)).cuda()
    obs, pi, z = (t.cpu().numpy() for t in replay.examples(idx))
    k = idx.cpu().numpy()
    assert np.array_equal(obs, want_obs[k]) and np.array_equal(pi, want_pi[k]) and np.array_equal(z, want_z[k])
    # a draw's z is +-1e-4 and tau=0 rows are one-hot on the move played
    assert set(np.unique(np.abs(want_z))) <= {np.float32(1.0), np.float32(1e-4)}
    onehot = host_recs["tau_one"] == 0
    assert (pi.shape[1] == 81) and np.array_equal(want_pi[::8][onehot].argmax(1), host_recs["action"][onehot])


def test_full_size_iteration_invariants_bf16():
    """BASELINE config C3's width (4096 trees, bf16 production evaluator, games to completion — fewer sims to keep the
    test short): size-independent properties of the whole iteration, checked by replaying every recorded move through
    the rules hooks: consecutive plies of a game chain through tc_rules_step, visit counts live on legal moves only and
    add up to at least sims - 1, the move played was visited, the recorded outcome is the terminal status of the last
    position seen from each record's player, and exactly numEnvs games were played (coach.py:175)."""
    import torch
    from tacult_b200.network import TrunkNet
    envs, sims = 4096, 24
    torch.manual_seed(3)
    sd = {k: v.detach().numpy().copy() for k, v in TrunkNet(32, 3, 256, dropout=0.2).eval().state_dict().items()}
    eng = tb.Engine(envs, sims * 82 + 64)
    eng.load_weights(sd)
    eng.set_evaluator(tb._lib.EVAL_NET_BF16)
    eng.selfplay_begin(sims, 2.4, 10, seed=77, auto_reset=False)
    for _ in range(81):
        eng.selfplay_step(1)
        if eng.selfplay_stats()["games_finished"] == envs:
            break
    eng.selfplay_sync()
    st = eng.selfplay_stats()
    recs = eng.selfplay_drain()
    assert st["games_finished"] == envs and len(recs) == st["plies"] and st["sims"] == st["plies"] * sims
    assert st["first_player_wins"] + st["second_player_wins"] + st["draws"] == envs
    order = np.lexsort((recs["ply"], recs["env"]))
    recs = recs[order]
    env, ply = recs["env"], recs["ply"]
    first = np.r_[True, env[1:] != env[:-1]]
    last = np.r_[first[1:], True]
    assert sorted(set(env.tolist())) == list(range(envs)) and (ply[first] == 0).all()
    assert (ply[~first] == ply[np.flatnonzero(~first) - 1] + 1).all()                 # plies 0..n-1 of every game
    assert (recs["state"][first] == 0).all() and (recs["player"][first] == 1).all()   # initial position, player +1
    counts = recs["counts"].astype(np.int64)
    legal = tb.rules_legal_mask(np.ascontiguousarray(recs["state"])).astype(bool)
    assert (counts[~legal] == 0).all() and (counts.sum(1) >= sims - 1).all()
    act = recs["action"].astype(np.int64)
    assert (counts[np.arange(len(recs)), act] > 0).all()
    greedy = recs["tau_one"] == 0
    assert (counts[np.arange(len(recs)), act][greedy] == counts.max(1)[greedy]).all()   # tau = 0: an arg-max was played
    assert ((recs["tau_one"] == 1) == (ply + 1 < 10)).all()                              # coach.py:124-125
    nxt, status = tb.rules_step(np.ascontiguousarray(recs["state"]), recs["action"])
    inner = np.flatnonzero(~last)
    assert (status[inner] == 0).all() and (nxt[inner] == recs["state"][inner + 1]).all()
    assert (recs["player"][inner + 1] == -recs["player"][inner]).all()
    end = np.flatnonzero(last)
    assert (status[end] != 0).all()
    # outcome per game: status of the final position is seen by the player who did NOT make the last move
    mover = recs["player"][end].astype(np.int64)
    winner = np.where(status[end] == tb._lib.LOST, mover, np.where(status[end] == tb._lib.WON, -mover, 0))
    game_of = np.cumsum(first) - 1
    w = winner[game_of]
    want_sign = np.where(w == 0, np.where(recs["player"] == -mover[game_of], 1, -1), np.where(recs["player"] == w, 1, -1))
    assert (recs["z_sign"] == want_sign).all()
    assert ((recs["finished"] == 2) == (w == 0)).all() and (recs["finished"] >= 1).all()
    assert int((winner == 1).sum()) == st["first_player_wins"] and int((winner == 0).sum()) == st["draws"]
