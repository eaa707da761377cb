"""Oracle search restatement (oracle/mcts_oracle.py) against root visit counts recorded from the LIVE
reference VectorizedMCTS (tests/golden/mcts_*.npz, made by tests/golden/make_golden.py)."""
import logging

import numpy as np
import pytest

from evaluators import HashEvaluator
from game_oracle import OracleGame
from mcts_oracle import OracleVectorizedMCTS
from utac_gym.core import GameState

logging.disable(logging.ERROR)     # the zero-prior fallback logs an error per hit, like mcts.py:281


def replay_selfplay(g, make_args, max_plies):
    args = make_args(numEnvs=int(g["envs"]), numMCTSSims=int(g["sims"]), cpuct=float(g["cpuct"]))
    mcts = OracleVectorizedMCTS(OracleGame(), HashEvaluator(int(g["salt"]), int(g["pi_levels"]), int(g["v_levels"])), args)
    for ply in range(min(len(g["roots"]), max_plies)):
        boards = [GameState.from_packed(w) for w in g["roots"][ply]]
        for _ in range(args.numMCTSSims):
            mcts.search(boards)
        counts = mcts.root_counts(boards)
        assert (counts == g["counts"][ply]).all(), f"ply {ply}"
    return mcts


@pytest.mark.parametrize("name,max_plies", [("mcts_selfplay_c1.npz", 999), ("mcts_selfplay_a.npz", 14),
                                            ("mcts_selfplay_ties.npz", 10)])
def test_selfplay_counts_match_reference(golden, make_args, name, max_plies):
    mcts = replay_selfplay(golden(name), make_args, max_plies)
    assert mcts.stat_sims > 0 and mcts.stat_nn_rows >= mcts.stat_leaf_evals


@pytest.mark.parametrize("name", ["mcts_arena_2sims.npz", "mcts_arena_16sims.npz"])
def test_arena_counts_match_reference(golden, make_args, name):
    g = golden(name)
    args = make_args(numEnvs=int(g["envs"]), numMCTSSims=int(g["sims"]), cpuct=float(g["cpuct"]))
    engines = [OracleVectorizedMCTS(OracleGame(), HashEvaluator(int(s)), args) for s in g["salts"]]
    for ply in range(len(g["roots"])):
        m = engines[int(g["movers"][ply])]
        boards = [GameState.from_packed(w) for w in g["roots"][ply]]
        for _ in range(args.numMCTSSims):
            m.search(boards)
        assert (m.root_counts(boards) == g["counts"][ply]).all(), f"ply {ply}"


def test_action_probs_semantics(make_args):
    args = make_args(numEnvs=3, numMCTSSims=12, cpuct=2.4)
    mcts = OracleVectorizedMCTS(OracleGame(), HashEvaluator(3), args)
    boards = [GameState() for _ in range(3)]
    np.random.seed(0)
    probs = mcts.getActionProbs(boards, temps=np.array([1, 0, 1]))
    counts = mcts.root_counts(boards)
    assert probs.shape == (3, 81) and probs.dtype == np.float64
    assert np.allclose(probs[0], counts[0] / counts[0].sum()) and counts[0].sum() == 11   # first sim expands the root
    assert probs[1].sum() == 1 and probs[1].max() == 1 and counts[1][probs[1].argmax()] == counts[1].max()
