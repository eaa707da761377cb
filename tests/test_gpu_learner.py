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


def test_training_from_device_replay_equals_training_from_host_arrays():
    """Same seeds, same batches: the learner fed by `DeviceReplay` (records in HBM, examples built per step by
    tc_examples_from_records_dev) ends with the very weights of the learner fed by the materialised host arrays."""
    from tacult_b200.distributed import DeviceReplay, DeviceTrajectoryExchange
    args = Args(numEnvs=48, numMCTSSims=6, cpuct=2.4, tempThreshold=8, tc_evaluator="f32")
    torch.manual_seed(1)
    net = TrunkNet(16, 1, 64, dropout=0.0)
    driver = SelfPlayDriver(net, args)
    eng = driver.engine
    eng.selfplay_begin(6, 2.4, 8, seed=4, auto_reset=False)
    eng.selfplay_step(81)
    eng.selfplay_sync()
    records, counts = DeviceTrajectoryExchange(eng, "cuda:0").gather()
    replay = DeviceReplay(eng, records)
    host = records.cpu().numpy().reshape(-1).view(tb._lib.EXAMPLE_DTYPE)
    obs, pi, z = records_to_arrays(host, augment=True)
    assert replay.n_examples == len(z) > 1000
    # an id outside the record array gives a zero example, never a wild read
    bad = torch.tensor([0, replay.n_examples + 5, -3], dtype=torch.int32, device="cuda")
    bo, bp, bz = replay.examples(bad)
    assert torch.equal(bo[0].cpu(), torch.from_numpy(obs[0])) and float(bo[1:].abs().sum() + bp[1:].abs().sum() + bz[1:].abs().sum()) == 0.0
    results = []
    for source in ("host", "device"):
        torch.manual_seed(2)
        learner = Learner(TrunkNet(16, 1, 64, dropout=0.0), lr=1e-3, device="cuda:0")
        if source == "host":
            hist = learner.train(obs, pi, z, epochs=1, steps_per_epoch=6, batch_size=512, seed=9)
        else:
            hist = learner.train_from_replay(replay, epochs=1, steps_per_epoch=6, batch_size=512, seed=9)
        results.append((hist, {k: v.detach().cpu().clone() for k, v in learner.net.state_dict().items()}))
    # cuDNN's backward kernels are not bit-reproducible from run to run, and Adam turns float32 noise in a small
    # gradient into a step of the order of lr: same batches => same losses, weights within a few steps' worth of lr
    assert np.allclose(results[0][0], results[1][0], rtol=1e-3)
    for k in results[0][1]:
        assert torch.allclose(results[0][1][k].float(), results[1][1][k].float(), rtol=0, atol=6 * 1e-3 + 1e-6), k
