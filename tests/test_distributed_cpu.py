"""Multi-process host logic on CPU: gloo backend, world_size 2 (the N>1 path of SURVEY.md §8e)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacult_b200._lib import EXAMPLE_DTYPE
from tacult_b200.distributed import all_gather_records, all_reduce_gradients, broadcast_weights, shard_envs


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def make_records(rank, n):
    rec = np.zeros(n, dtype=EXAMPLE_DTYPE)
    rec["env"] = np.arange(n) + 1000 * rank
    rec["ply"] = np.arange(n) % 60
    rec["counts"][:, rank] = 7
    rec["state"][:, 0] = np.arange(n) * 3 + rank
    rec["z_sign"] = 1 if rank == 0 else -1
    rec["finished"] = 1
    return rec


def worker(rank, world, port, counts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = all_gather_records(make_records(rank, counts[rank]))
        want = np.concatenate([make_records(r, counts[r]) for r in range(world)])
        assert got.dtype == EXAMPLE_DTYPE and len(got) == sum(counts)
        for name in EXAMPLE_DTYPE.names:          # field by field: numpy does not define the padding bytes
            assert np.array_equal(got[name], want[name]), name
        # learner-side collectives
        torch.manual_seed(rank)
        net = torch.nn.Linear(4, 3)
        broadcast_weights(net, src=0)
        ref = torch.nn.Linear(4, 3)
        torch.manual_seed(0)
        ref = torch.nn.Linear(4, 3)
        assert torch.equal(net.weight, ref.weight) and torch.equal(net.bias, ref.bias)
        x = torch.full((2, 4), float(rank + 1))
        net(x).sum().backward()
        all_reduce_gradients(net)
        expect = torch.full((3, 4), 2.0 * (1 + 2) / 2)        # mean over ranks of sum_x x
        assert torch.allclose(net.weight.grad, expect)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(5, 9), (0, 4), (3, 0), (0, 0)])
def test_all_gather_records_gloo_world2(tmp_path, counts):
    port = free_port()
    mp.spawn(worker, args=(2, port, counts, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_envs_covers_every_env_once():
    for total, world in [(32768, 8), (4096, 1), (10, 4), (7, 8)]:
        blocks = [shard_envs(total, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert shard_envs(32768, 8, 3) == (3 * 4096, 4 * 4096)
