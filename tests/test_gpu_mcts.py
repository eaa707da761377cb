"""CUDA search engine vs the reference: root visit counts must be IDENTICAL when both sides see identical
evaluator outputs (BASELINE.json north_star).  Three anchors:
  1. golden counts recorded from the live reference VectorizedMCTS (tests/golden/mcts_*.npz);
  2. the oracle restatement run here on fresh seeds (also compares every root Q bit for bit);
  3. record/replay: the GPU engine's own fp32-network leaf evaluations replayed through the oracle search."""
import logging

import numpy as np
import pytest

import resnet_oracle
import tacult_b200 as tb
from evaluators import HashEvaluator, ReplayEvaluator
from game_oracle import OracleGame
from mcts_oracle import OracleVectorizedMCTS
from utac_gym.core import GameState

pytestmark = pytest.mark.gpu
logging.disable(logging.ERROR)


class HashNet:
    """Tells tacult_b200.VectorizedMCTS to use the device-side hash evaluator (same spec as HashEvaluator)."""
    def __init__(self, salt, pi_levels=0, v_levels=0):
        self.tc_hash_params = (salt, pi_levels, v_levels)


@pytest.mark.parametrize("name", ["mcts_selfplay_a.npz", "mcts_selfplay_ties.npz", "mcts_selfplay_c1.npz"])
def test_selfplay_counts_match_reference_golden(golden, make_args, name):
    g = golden(name)
    args = make_args(numEnvs=int(g["envs"]), numMCTSSims=int(g["sims"]), cpuct=float(g["cpuct"]))
    mcts = tb.VectorizedMCTS(OracleGame(), HashNet(int(g["salt"]), int(g["pi_levels"]), int(g["v_levels"])), args)
    for ply in range(len(g["roots"])):
        boards = [GameState.from_packed(w) for w in g["roots"][ply]]
        counts = mcts.rootCounts(boards)
        assert (counts == g["counts"][ply]).all(), f"{name} ply {ply}"
    st = mcts.engine.stats()
    assert st["sims"] == int(g["alive"].sum()) * int(g["sims"])


@pytest.mark.parametrize("name", ["mcts_arena_2sims.npz", "mcts_arena_16sims.npz"])
def test_arena_counts_match_reference_golden(golden, make_args, name):
    g = golden(name)
    args = make_args(numEnvs=int(g["envs"]), numMCTSSims=int(g["sims"]), cpuct=float(g["cpuct"]))
    engines = [tb.VectorizedMCTS(OracleGame(), HashNet(int(s)), args) for s in g["salts"]]
    for ply in range(len(g["roots"])):
        boards = [GameState.from_packed(w) for w in g["roots"][ply]]
        counts = engines[int(g["movers"][ply])].rootCounts(boards)
        assert (counts == g["counts"][ply]).all(), f"{name} ply {ply}"


def play_both(make_args, envs, sims, cpuct, salt, pi_levels, v_levels, seed, max_plies=200):
    """GPU engine and oracle search side by side on a fresh self-play; counts and root Q compared every ply."""
    args = make_args(numEnvs=envs, numMCTSSims=sims, cpuct=cpuct)
    game = OracleGame()
    gpu = tb.VectorizedMCTS(game, HashNet(salt, pi_levels, v_levels), args)
    cpu = OracleVectorizedMCTS(game, HashEvaluator(salt, pi_levels, v_levels), args)
    boards = [game.getInitBoard() for _ in range(envs)]
    players = np.ones(envs, dtype=int)
    rng = np.random.RandomState(seed)
    plies = 0
    while plies < max_plies:
        alive = [game.getGameEnded(boards[i], players[i]) == 0 for i in range(envs)]
        if not any(alive):
            break
        canon = [game.getCanonicalForm(boards[i], players[i]) for i in range(envs)]
        got = gpu.rootCounts(canon)
        for _ in range(sims):
            cpu.search(canon)
        want = cpu.root_counts(canon)
        assert (got == want).all(), f"ply {plies}"
        q = gpu.engine.root_q()
        for i in range(envs):
            key = game.stringRepresentation(canon[i])
            for a in range(81):
                ref_q = cpu.Qsa[i].get((key, a))
                if ref_q is None:
                    assert np.isnan(q[i, a])
                else:
                    assert q[i, a] == ref_q, (plies, i, a)      # float64, bit for bit
        for i in range(envs):
            if alive[i]:
                p = want[i].astype(float)
                boards[i], players[i] = game.getNextState(boards[i], int(players[i]), int(rng.choice(81, p=p / p.sum())))
        plies += 1
    return gpu, cpu


def test_counts_and_q_match_oracle_fresh_seeds(make_args):
    gpu, cpu = play_both(make_args, envs=5, sims=48, cpuct=2.4, salt=1234, pi_levels=0, v_levels=0, seed=1)
    st = gpu.engine.stats()
    assert st["sims"] == cpu.stat_sims - 0 or st["sims"] <= cpu.stat_sims     # oracle also counts sims on finished roots
    assert st["nn_leaves"] == cpu.stat_leaf_evals
    assert st["sum_depth"] == cpu.stat_sum_depth
    assert st["nodes_in_use"] == sum(len(x) for x in cpu.Ps)


def test_counts_match_oracle_with_ties_zero_priors_and_draw_values(make_args):
    play_both(make_args, envs=4, sims=32, cpuct=1.25, salt=77, pi_levels=2, v_levels=1, seed=2)


def test_deep_search_transpositions(make_args):
    # a configuration where the reference search merges transposed positions (found with the oracle:
    # first use of an edge whose successor is already a node); the engine must merge the same ones
    gpu, cpu = play_both(make_args, envs=2, sims=300, cpuct=4.0, salt=6, pi_levels=2, v_levels=0, seed=3, max_plies=12)
    assert gpu.engine.stats()["transpositions"] == cpu.stat_transpositions > 0


def test_none_and_terminal_roots(make_args, golden):
    args = make_args(numEnvs=4, numMCTSSims=10, cpuct=2.4)
    mcts = tb.VectorizedMCTS(OracleGame(), HashNet(3), args)
    fin = GameState.from_packed(golden("rules_playouts.npz")["final_state"][0])
    assert fin.is_terminal()
    counts = mcts.rootCounts([GameState(), None, fin, GameState()])
    assert counts[0].sum() == 9 and counts[3].sum() == 9 and counts[1].sum() == 0 and counts[2].sum() == 0
    np.random.seed(0)
    probs = mcts.getActionProbs([GameState(), None, fin, GameState()], temps=np.array([1, 1, 0, 0]))
    assert abs(probs[0].sum() - 1) < 1e-12 and np.isnan(probs[1]).all() and probs[2].sum() == 1 and probs[3].max() == 1


def test_reset_and_capacity_error(make_args):
    args = make_args(numEnvs=2, numMCTSSims=20, cpuct=2.4, tc_max_nodes_per_tree=8)
    mcts = tb.VectorizedMCTS(OracleGame(), HashNet(3), args)
    with pytest.raises(tb._lib.TacultError) as err:
        mcts.rootCounts([GameState(), GameState()])
    assert err.value.code == tb._lib.TC_ECAPACITY
    args = make_args(numEnvs=2, numMCTSSims=20, cpuct=2.4)
    mcts = tb.VectorizedMCTS(OracleGame(), HashNet(3), args)
    first = mcts.rootCounts([GameState(), GameState()])
    again = mcts.rootCounts([GameState(), GameState()])
    assert again.sum() == first.sum() + 40            # statistics persist (coach.py:162,175-176)
    mcts.reset(0)
    third = mcts.rootCounts([GameState(), GameState()])
    assert (third[0] == first[0]).all() and third[1].sum() == again[1].sum() + 20


def test_full_size_properties():
    """C3-sized engine (4096 trees): size-independent invariants of one search call."""
    eng = tb.Engine(4096, 200 * 4 + 64, hash_salt=9)
    eng.set_roots_packed(np.zeros((4096, 8), dtype=np.uint32))
    eng.search(200, 2.4)
    counts = eng.root_counts()
    assert (counts.sum(1) == 199).all()               # the first simulation expands the root
    assert (counts == counts[0]).all()                # identical positions + identical evaluator => identical trees
    st = eng.stats()
    assert st["sims"] == 4096 * 200 and st["nn_leaves"] + st["terminal_leaves"] == st["sims"]
    assert st["max_nodes_one_tree"] <= 200


@pytest.mark.parametrize("parts", ["2", "3", "4"])
def test_partitioned_engine_gives_the_same_counts(parts, monkeypatch):
    """TACULT_HALVES splits the trees over streams (read at tc_create); the trees are independent, so every root count
    and Q must equal the single-partition engine's, for different positions per tree and a ragged last partition."""
    n = 600
    rng = np.random.RandomState(11)
    res = tb.rules_playouts(5, 0, n, max_plies=12)           # distinct mid-game roots
    roots = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    active = (rng.random_sample(n) < 0.9).astype(np.uint8)

    def run():
        eng = tb.Engine(n, 40 * 4 + 64, hash_salt=17)
        eng.set_roots_packed(roots, active)
        eng.search(40, 2.4)
        return eng.root_counts(), eng.root_q()

    monkeypatch.setenv("TACULT_HALVES", "1")
    c1, q1 = run()
    monkeypatch.setenv("TACULT_HALVES", parts)
    c2, q2 = run()
    assert np.array_equal(c1, c2) and np.array_equal(q1.view(np.uint64), q2.view(np.uint64))
    assert (c1[active == 0] == 0).all() and c1[active == 1].sum() > 0


@pytest.mark.parametrize("mode,arch,tol", [("f32", (16, 2, 64), 1e-3), ("bf16", (32, 3, 256), 2e-2)])
def test_record_replay_with_network_evaluator(make_args, mode, arch, tol):
    """Identical evaluator outputs, real network: the engine logs every leaf (state, pi, v) its ResNet produced — the
    fp32 CUDA-core chain or the production bf16 chain (fused tcgen05 trunk -> split-K fc1 -> tensor-core heads); the
    oracle search replays those exact float32 values and must reach the same visit counts.  This pins the search
    bookkeeping around the production evaluator (leaf compaction, row routing, counter pair), not only its forward."""
    envs, sims = 6, 40
    if mode == "bf16":
        # the 2e-2 bar of the north_star is stated for the reference's own network: torch-default initialisation of
        # `_UtacNNet` (network.py:84-101), which is what TrunkNet reproduces.  On the He-scaled synthetic weights of
        # make_state_dict (policy head gain 2, logits of +-6) ANY bf16 chain has a longer tail: 1.1e-4 mean, 1e-2 at
        # the 99.9th percentile, 3.8e-2 maximum over 2048 positions, for the tap-per-UMMA trunk and the column-tiled one
        # alike (tools/probes/bf16_error_ab.py).
        import torch
        from tacult_b200.network import TrunkNet
        torch.manual_seed(11)
        sd = {k: v.detach().numpy().copy() for k, v in TrunkNet(arch[0], arch[1], arch[2], dropout=0.2).eval().state_dict().items()}
    else:
        sd = resnet_oracle.make_state_dict(11, channels=arch[0], num_residual_blocks=arch[1], embedding_size=arch[2])

    class Net:                     # the shape tacult_b200.VectorizedMCTS expects: .nnet.state_dict()
        class nnet:
            @staticmethod
            def state_dict():
                return sd
    args = make_args(numEnvs=envs, numMCTSSims=sims, cpuct=2.4, tc_eval_log_capacity=envs * sims * 20, tc_evaluator=mode)
    game = OracleGame()
    gpu = tb.VectorizedMCTS(game, Net(), args)
    boards = [game.getInitBoard() for _ in range(envs)]
    players = np.ones(envs, dtype=int)
    rng = np.random.RandomState(4)
    history = []
    for ply in range(16):
        canon = [game.getCanonicalForm(boards[i], players[i]) for i in range(envs)]
        counts = gpu.rootCounts(canon)
        history.append((canon, counts))
        for i in range(envs):
            p = counts[i].astype(float)
            boards[i], players[i] = game.getNextState(boards[i], int(players[i]), int(rng.choice(81, p=p / p.sum())))
    states, pi, v, total = gpu.engine.eval_log(envs * sims * 20)
    assert total == len(states) > 0
    table = {}
    for k in range(len(states)):
        obs = np.asarray(GameState.from_packed(states[k]).get_obs())
        table[ReplayEvaluator.key_of(obs)] = (pi[k], v[k])
    cpu = OracleVectorizedMCTS(game, ReplayEvaluator(table), make_args(numEnvs=envs, numMCTSSims=sims, cpuct=2.4))
    for canon, counts in history:
        for _ in range(sims):
            cpu.search(canon)
        assert (cpu.root_counts(canon) == counts).all()
    # and the logged evaluations themselves are the reference network's, to the mode's tolerance
    obs = np.stack([np.asarray(GameState.from_packed(s).get_obs()).reshape(4, 9, 9) for s in states[:256]])
    rpi, rv = resnet_oracle.forward(sd, obs)
    assert np.abs(rpi.numpy() - pi[:256]).max() < tol and np.abs(rv.numpy().ravel() - v[:256]).max() < tol


def test_set_roots_fields_form_equals_packed_form():
    """tc_set_roots takes the GAMESTATE members (uint16 board[9], occ[9], int8 last_square) as the reference holds them
    (utac_game.py:101-106); it must plant the same roots as tc_set_roots_packed."""
    n = 96
    res = tb.rules_playouts(21, 0, n, max_plies=25)
    roots = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    board = np.zeros((n, 9), dtype=np.uint16)
    occ = np.zeros((n, 9), dtype=np.uint16)
    for i in range(9):
        board[:, i] = (roots[:, i // 3] >> (9 * (i % 3))) & 0x1FF
        occ[:, i] = (roots[:, 3 + i // 3] >> (9 * (i % 3))) & 0x1FF
    last = (roots[:, 6] & 0xF).astype(np.int8) - 1
    active = (np.arange(n) % 5 != 0).astype(np.uint8)
    a = tb.Engine(n, 30 * 4 + 64, hash_salt=3)
    a.set_roots_packed(roots, active)
    a.search(30, 2.4)
    b = tb.Engine(n, 30 * 4 + 64, hash_salt=3)
    b.set_roots_fields(board, occ, last, active)
    b.search(30, 2.4)
    ca, cb = a.root_counts(), b.root_counts()
    assert np.array_equal(ca, cb) and np.array_equal(a.root_q().view(np.uint64), b.root_q().view(np.uint64))
    live = tb.rules_status(roots) == 0
    assert (ca[(active == 1) & live].sum(1) == 29).all() and (ca[active == 0] == 0).all()


def test_reset_envs_equals_per_env_resets():
    """tc_reset_envs(list) == VectorizedMCTS.reset(i) for every i of the list (mcts.py:165-172), in one launch."""
    n = 40
    roots = np.zeros((n, 8), dtype=np.uint32)
    engines = []
    for batched in (False, True):
        eng = tb.Engine(n, 30 * 4 + 64, hash_salt=21)
        eng.set_roots_packed(roots)
        eng.search(20, 2.4)
        which = np.array([0, 3, 17, 39])
        if batched:
            eng.reset_envs(which)
        else:
            for i in which:
                eng.reset(int(i))
        eng.set_roots_packed(roots)
        eng.search(10, 2.4)
        engines.append((eng.root_counts(), eng.root_q()))
    (c0, q0), (c1, q1) = engines
    assert np.array_equal(c0, c1) and np.array_equal(q0.view(np.uint64), q1.view(np.uint64))
    assert (c0[[0, 3, 17, 39]].sum(1) == 9).all() and (np.delete(c0, [0, 3, 17, 39], axis=0).sum(1) == 29).all()
    with pytest.raises(tb._lib.TacultError):
        tb.Engine(4, 64).reset_envs([7])


def test_advance_on_device_equals_host_stepping():
    """tc_advance (root + action -> next root on the device, trees kept, finished games restarted) against the host-driven
    sequence it replaces: tc_rules_step + tc_reset_envs + tc_set_roots_packed.  Same visit counts and Q at every ply."""
    n, sims = 96, 24
    rng = np.random.RandomState(5)
    a = tb.Engine(n, sims * 82 + 64, hash_salt=33)
    b = tb.Engine(n, sims * 82 + 64, hash_salt=33)
    states = np.zeros((n, 8), dtype=np.uint32)
    a.set_roots_packed(states)
    b.set_roots_packed(states)
    finished = 0
    for ply in range(60):
        a.search(sims, 2.4)
        b.search(sims, 2.4)
        ca, cb = a.root_counts(), b.root_counts()
        assert np.array_equal(ca, cb) and np.array_equal(a.root_q().view(np.uint64), b.root_q().view(np.uint64)), ply
        acts = np.array([rng.choice(np.flatnonzero(c)) for c in ca], dtype=np.int8)
        # host path
        nxt, st = tb.rules_step(states, acts)
        done = np.flatnonzero(st != 0)
        nxt[done] = 0
        a.reset_envs(done)
        a.set_roots_packed(nxt)
        # device path
        got, gst = b.advance(acts, restart=True)
        assert np.array_equal(gst, st)
        live = st == 0
        assert np.array_equal(got[live], nxt[live])
        states = nxt
        finished += len(done)
    assert finished >= 3                       # restarts were exercised
    # an illegal action is refused loudly and leaves that env alone
    bad = np.full(n, -1, dtype=np.int8)
    occupied = int(np.flatnonzero(tb.rules_legal_mask(states[:1])[0] == 0)[0])
    bad[0] = occupied
    with pytest.raises(tb._lib.TacultError) as err:
        b.advance(bad)
    assert err.value.code == tb._lib.TC_EILLEGAL
