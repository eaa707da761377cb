"""Learner step (tacult_b200/learner.py) against the LIVE reference learner recorded in tests/golden/learner_steps.npz
(three NNetWrapper.training_step calls with drop-out active, nn_wrapper.py:45-56), the data-parallel step on gloo
(world 2) against the single-process step, and checkpoint interchange with a file the reference wrote."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacult_b200.learner import Learner, learner_from_checkpoint, load_reference_checkpoint
from tacult_b200.network import TrunkNet

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-6          # same torch ops in the same order on the same build: in practice bit-identical


def golden_state(g, prefix):
    return {str(k): g[f"{prefix}/{k}"] for k in g["keys"]}


def test_three_steps_match_the_reference_learner(golden):
    g = golden("learner_steps.npz")
    learner = Learner(TrunkNet.from_state_dict(golden_state(g, "init"), dropout=0.2), lr=1e-3)
    learner.net.train()
    torch.manual_seed(5)
    for i in range(3):
        pi_loss, v_loss = learner.training_step(torch.from_numpy(g["obs"][i]), torch.from_numpy(g["pis"][i]),
                                                torch.from_numpy(g["vs"][i]))
        assert abs(pi_loss - g["losses"][i][0]) < TOL and abs(v_loss - g["losses"][i][1]) < TOL
    final = golden_state(g, "final")
    for k, v in learner.net.state_dict().items():
        assert np.allclose(v.numpy(), final[k], atol=TOL, rtol=0), k
    learner.net.eval()
    with torch.no_grad():
        pi, v = learner.net(torch.from_numpy(g["obs"][0]))
    assert np.allclose(pi.numpy(), g["eval_pi"], atol=TOL) and np.allclose(v.numpy(), g["eval_v"], atol=TOL)


def test_reference_checkpoint_loads_and_round_trips(golden, tmp_path):
    g = golden("learner_steps.npz")
    path = os.path.join(HERE, "golden", "ref_checkpoint_tiny.pt")
    sd, opt = load_reference_checkpoint(path)
    final = golden_state(g, "final")
    assert list(sd.keys()) == list(final.keys()) and opt is not None
    for k in sd:
        assert np.array_equal(sd[k], final[k]), k
    learner = learner_from_checkpoint(path, lr=1e-3, dropout=0.2)
    assert learner.optimizer.state_dict()["state"][0]["step"] == 3          # Adam moments came along
    learner.save_checkpoint(str(tmp_path), "again.pt")
    sd2, opt2 = load_reference_checkpoint(os.path.join(str(tmp_path), "again.pt"))
    assert all(np.array_equal(sd2[k], sd[k]) for k in sd) and set(opt2.keys()) == set(opt.keys())
    with np.testing.assert_raises(FileNotFoundError):
        learner.load_checkpoint(str(tmp_path), "missing.pt")


def test_fixed_step_epoch_loop_lowers_the_loss():
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    obs = (rng.random_sample((256, 4, 9, 9)) < 0.3).astype(np.float32)
    pi = np.zeros((256, 81), np.float32)
    pi[np.arange(256), obs[:, 0].reshape(256, 81).argmax(1)] = 1.0          # a learnable target
    z = np.where(obs[:, 1].sum((1, 2)) > 24, 1.0, -1.0).astype(np.float32)
    learner = Learner(TrunkNet(8, 1, 32, dropout=0.0), lr=3e-3)
    hist = learner.train(obs, pi, z, epochs=3, steps_per_epoch=20, batch_size=64, seed=1)
    assert learner.steps == 60 and hist[-1][0] < hist[0][0] and hist[-1][1] < hist[0][1]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def dp_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(3)
        init = TrunkNet(8, 1, 32, dropout=0.0).state_dict()
        rng = np.random.RandomState(4)
        obs = torch.from_numpy((rng.random_sample((24, 4, 9, 9)) < 0.3).astype(np.float32))
        pis = torch.from_numpy(rng.dirichlet(np.ones(81), 24).astype(np.float32))
        vs = torch.from_numpy(rng.choice([-1.0, 1.0], 24).astype(np.float32))

        def fresh(distributed):
            net = TrunkNet(8, 1, 32, dropout=0.0)
            net.load_state_dict(init)
            net.eval()                       # BatchNorm on running statistics: the sharded step is then exactly the full one
            return Learner(net, lr=1e-3, distributed=distributed)

        single = fresh(False)
        want = [single.training_step(obs, pis, vs) for _ in range(2)]
        sharded = fresh(True)
        rows = slice(0, 10) if rank == 0 else slice(10, 24)          # ragged shards
        got = [sharded.training_step(obs[rows], pis[rows], vs[rows]) for _ in range(2)]
        assert np.allclose(got, want, atol=1e-5), (got, want)
        for (k, a), b in zip(sharded.net.state_dict().items(), single.net.state_dict().values()):
            assert torch.allclose(a, b, atol=1e-5), k
        # diverge rank 1, then broadcast rank 0's weights
        if rank == 1:
            with torch.no_grad():
                sharded.net.pi.bias.add_(1.0)
        sharded.broadcast(src=0)
        assert torch.allclose(sharded.net.pi.bias, single.net.pi.bias, atol=1e-5)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_equals_single_process_step(tmp_path):
    mp.spawn(dp_worker, args=(2, free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(2))
