"""Run as a script (tests/test_gpu_callers.py spawns it): the UNMODIFIED reference `Coach.generateTrainExamples`
(staged copy under oracle/_ref, or /root/reference in the build container) and the reference's own `UtacGame` adapter,
with the two replacements of INTEGRATION.md and nothing of oracle/ on the rules path:

    utac_gym.core.GameState  <-  tacult_b200.GameState        (install_as_utac_gym)
    base.coach.MCTS          <-  tacult_b200.VectorizedMCTS   (the one-import change)

Prints the number of training examples and the sha256 of (obs, pi, z) over all of them; the caller compares with
tests/golden/callers.npz, recorded from the live reference over the oracle rules and the reference search."""
import contextlib
import hashlib
import io
import logging
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import tacult_b200 as tb  # noqa: E402

assert tb.install_as_utac_gym(force=True)            # before anything imports utac_gym
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402

logging.disable(logging.ERROR)
ref = ref_loader.load_reference()
from utac_gym.core import GameState  # noqa: E402
assert GameState is tb.GameState and ref.utac_game.GameState is tb.GameState


class HashNet:
    tc_hash_params = (5, 0, 0)


ref.coach.MCTS = tb.VectorizedMCTS                    # `from tacult_b200 import VectorizedMCTS as MCTS`
args = ref.utils.dotdict(numEnvs=4, minNumEps=4, numMCTSSims=6, cpuct=2.4, tempThreshold=5)
coach = ref.coach.Coach.__new__(ref.coach.Coach)
coach.game, coach.args, coach.pnet = ref.utac_game.UtacGame(), args, HashNet()
np.random.seed(7)
with contextlib.redirect_stdout(io.StringIO()):
    examples = coach.generateTrainExamples()
h = hashlib.sha256()
for obs, pi, z in examples:
    h.update(obs.numpy().tobytes())
    h.update(pi.numpy().tobytes())
    h.update(np.float64(z).tobytes())
print("EXAMPLES", len(examples), h.hexdigest())
