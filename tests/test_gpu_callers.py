"""The drop-in `tacult_b200.VectorizedMCTS` under the reference's callers: arena results and the iteration's
training examples must equal those the reference produces (tests/golden/callers.npz) — same visit counts, same
global-RNG draws (mcts.py:207-208, coach.py:140)."""
import logging

import numpy as np
import pytest

import tacult_b200 as tb
from test_oracle_callers import arena_with, examples_digest, selfplay_with

pytestmark = pytest.mark.gpu
logging.disable(logging.ERROR)


class HashNet:
    def __init__(self, salt):
        self.tc_hash_params = (salt, 0, 0)


def gpu_mcts(game, salt, args):
    return tb.VectorizedMCTS(game, HashNet(salt), args)


def test_arena_results_with_gpu_engine(golden, make_args):
    for seed, one, two, draws in golden("callers.npz")["arena"]:
        assert arena_with(gpu_mcts, make_args, int(seed)) == (one, two, draws)


def test_selfplay_examples_with_gpu_engine(golden, make_args):
    g = golden("callers.npz")
    ex = selfplay_with(gpu_mcts, make_args)
    assert len(ex) == int(g["coach_examples"])
    assert (examples_digest(ex) == g["coach_sha256"]).all()
