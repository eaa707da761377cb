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


def test_packed_arena_matches_reference_results(golden):
    """The native two-engine arena on packed positions (tacult_b200/arena.py) against the results the LIVE reference
    arena produced with the same evaluators and seeds: same counts -> same RNG draws -> same games."""
    from tacult_b200.arena import PackedArena
    for seed, one, two, draws in golden("callers.npz")["arena"]:
        np.random.seed(int(seed))
        e1 = tb.Engine(6, 4096)
        e2 = tb.Engine(6, 4096)
        e1.set_hash_evaluator(1)
        e2.set_hash_evaluator(2)
        arena = PackedArena(e1, e2, num_sims=3, cpuct=2.4)
        assert arena.play_games(20) == (one, two, draws)
        assert arena.sims > 0 and len(arena.ply_seconds) > 0


def test_packed_arena_lowest_index_mode_is_deterministic():
    from tacult_b200.arena import PackedArena
    results = []
    for _ in range(2):
        e1, e2 = tb.Engine(32, 4096), tb.Engine(32, 4096)
        e1.set_hash_evaluator(3)
        e2.set_hash_evaluator(4)
        results.append(PackedArena(e1, e2, num_sims=8, cpuct=2.4, tie_break="lowest").play_games(64))
    assert results[0] == results[1] and sum(results[0]) == 64


def test_reference_coach_runs_on_product_gamestate_and_gpu_search(golden):
    """INTEGRATION.md end to end: the unmodified reference `Coach.generateTrainExamples` + `UtacGame` with
    `utac_gym.core.GameState` served by tacult_b200.GameState and `MCTS` replaced by tacult_b200.VectorizedMCTS gives
    the training examples the live reference produced (callers.npz), bit for bit.  Needs the staged reference
    modules (oracle/_ref, written by __graft_entry__.build() in the build container)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if not (os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "src", "tacult", "base", "coach.py"))
            or os.path.isfile("/root/reference/src/tacult/base/coach.py")):
        pytest.skip("reference modules not staged (oracle/_ref)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_reference_coach.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("EXAMPLES")][-1].split()
    g = golden("callers.npz")
    assert int(line[1]) == int(g["coach_examples"])
    assert line[2] == bytes(g["coach_sha256"]).hex()
