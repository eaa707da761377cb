"""Arena and self-play-loop restatements (oracle/callers_oracle.py) against results recorded from the LIVE
reference `VectorizedArena.playGames` and `Coach.generateTrainExamples` (tests/golden/callers.npz)."""
import hashlib
import logging

import numpy as np

from callers_oracle import arena_play_games, selfplay_iteration
from evaluators import HashEvaluator
from game_oracle import OracleGame
from mcts_oracle import OracleVectorizedMCTS

logging.disable(logging.ERROR)


def arena_with(make_mcts, make_args, seed):
    args = make_args(numEnvs=6, numMCTSSims=3, cpuct=2.4)
    np.random.seed(seed)
    game = OracleGame()
    m1, m2 = make_mcts(game, 1, args), make_mcts(game, 2, args)
    return arena_play_games(game, lambda x: np.argmax(m1.getActionProbs(x, temps=np.zeros(6)), axis=1),
                            lambda x: np.argmax(m2.getActionProbs(x, temps=np.zeros(6)), axis=1), 20, 6)


def examples_digest(examples):
    h = hashlib.sha256()
    for obs, pi, z in examples:
        h.update(obs.numpy().tobytes())
        h.update(pi.numpy().tobytes())
        h.update(np.float64(z).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def selfplay_with(make_mcts, make_args):
    args = make_args(numEnvs=4, minNumEps=4, numMCTSSims=6, cpuct=2.4, tempThreshold=5)
    np.random.seed(7)
    game = OracleGame()
    return selfplay_iteration(game, make_mcts(game, 5, args), args)


def oracle_mcts(game, salt, args):
    return OracleVectorizedMCTS(game, HashEvaluator(salt), args)


def test_arena_results_match_reference(golden, make_args):
    for seed, one, two, draws in golden("callers.npz")["arena"]:
        assert arena_with(oracle_mcts, make_args, int(seed)) == (one, two, draws)


def test_selfplay_examples_match_reference(golden, make_args):
    g = golden("callers.npz")
    ex = selfplay_with(oracle_mcts, make_args)
    assert len(ex) == int(g["coach_examples"])
    assert (examples_digest(ex) == g["coach_sha256"]).all()
