"""bf16 chain vs the float32 oracle on the synthetic He-scaled net of tests/test_gpu_mcts.py: error statistics of the
forward for the trunk selected by TACULT_TRUNK_COL (run once with 0 and once with 1)."""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import resnet_oracle  # noqa: E402
import tacult_b200 as tb  # noqa: E402
from utac_gym.core import GameState  # noqa: E402  (oracle rules: test infrastructure)


def positions(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        s = GameState()
        for _ in range(rng.randint(0, 50)):
            if s.is_terminal():
                break
            s.make_move(int(rng.choice(np.flatnonzero(s.get_valid_mask()))))
        if not s.is_terminal():
            out.append(np.asarray(s.get_obs(), dtype=np.float32).reshape(4, 9, 9))
    return np.stack(out)


def main():
    for seed in (11, 12, 13):
        sd = resnet_oracle.make_state_dict(seed, channels=32, num_residual_blocks=3, embedding_size=256)
        obs = positions(2048, seed)
        eng = tb.Engine(64, 256)
        eng.load_weights(sd)
        pi, v = eng.forward(obs, tb._lib.EVAL_NET_BF16)
        rpi, rv = resnet_oracle.forward(sd, obs)
        dpi = np.abs(rpi.numpy() - pi)
        dv = np.abs(rv.numpy().ravel() - v)
        print(f"TRUNK_COL={os.environ.get('TACULT_TRUNK_COL', '1')} seed {seed}: pi max {dpi.max():.5f} mean {dpi.mean():.6f} "
              f"p99.9 {np.quantile(dpi, 0.999):.5f} | v max {dv.max():.5f} mean {dv.mean():.6f}")


if __name__ == "__main__":
    main()
