"""Actor -> learner -> actor on the GPU: device-resident self-play produces examples, the learner
(tacult_b200/learner.py) takes optimisation steps on them, the new weights go back into the engine and the
engine's forward must agree with the learner's eval-mode network."""
import numpy as np
import pytest
import torch

import tacult_b200 as tb
from tacult_b200.learner import Learner
from tacult_b200.network import TrunkNet
from tacult_b200.selfplay import SelfPlayDriver, records_to_arrays

pytestmark = pytest.mark.gpu


class Args(dict):
    __getattr__ = dict.__getitem__


def test_selfplay_train_publish_round_trip():
    torch.manual_seed(0)
    net = TrunkNet(32, 3, 256, dropout=0.2)
    args = Args(numEnvs=64, numMCTSSims=8, cpuct=2.4, tempThreshold=10, tc_evaluator="f32")
    driver = SelfPlayDriver(net, args)
    records, stats = driver.play_records(seed=1)
    assert stats["games_finished"] == 64
    obs, pi, z = records_to_arrays(records, augment=True)
    assert len(z) == 8 * len(records)

    learner = Learner(net, lr=1e-3, device="cuda:0")
    before = {k: v.clone() for k, v in learner.net.state_dict().items()}
    hist = learner.train(obs, pi, z, epochs=2, steps_per_epoch=10, batch_size=256, seed=2)
    assert learner.steps == 20 and all(np.isfinite(h).all() for h in hist)
    assert any(not torch.equal(before[k], v) for k, v in learner.net.state_dict().items())

    learner.publish(driver.engine)                      # tc_load_weights: BatchNorm (with its new running stats) re-folded
    learner.net.eval()
    with torch.no_grad():
        want_pi, want_v = learner.net(torch.from_numpy(obs[:128]).cuda())
    got_pi, got_v = driver.engine.forward(obs[:128], tb._lib.EVAL_NET_F32)
    assert np.abs(got_pi - want_pi.cpu().numpy()).max() < 1e-3 and np.abs(got_v - want_v.cpu().numpy().ravel()).max() < 1e-3
    got_pi, got_v = driver.engine.forward(obs[:128], tb._lib.EVAL_NET_BF16)
    assert np.abs(got_pi - want_pi.cpu().numpy()).max() < 2e-2 and np.abs(got_v - want_v.cpu().numpy().ravel()).max() < 2e-2
    # and the engine keeps playing with the new weights
    records2, stats2 = driver.play_records(seed=3)
    assert stats2["games_finished"] >= 64 and len(records2) > 0
