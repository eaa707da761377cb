"""NCCL path of the one exchange step, world_size 2 on one box: every rank plays its own shard of games,
drains the records ON THE DEVICE (tc_selfplay_drain_dev), all-gathers them over NCCL (DeviceTrajectoryExchange) and must
end up with, in rank order, exactly what each rank's host drain (tc_selfplay_drain) of an identical engine returns;
then a data-parallel learner step straight from the device replay.  Skipped on a single-GPU box (run with
`gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def play(tb, envs, sims, seed, device):
    eng = tb.Engine(envs, sims * 82 + 64, hash_salt=13, device=device)
    eng.selfplay_begin(sims, 2.4, 8, seed=seed, auto_reset=False)
    eng.selfplay_step(81)
    eng.selfplay_sync()
    return eng


def worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import tacult_b200 as tb
    from tacult_b200.distributed import DeviceReplay, DeviceTrajectoryExchange
    from tacult_b200.learner import Learner
    from tacult_b200.network import TrunkNet
    from tacult_b200.selfplay import records_to_arrays
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        envs, sims = 48 + 16 * rank, 10              # ragged: the ranks hold different record counts
        eng = play(tb, envs, sims, 100 + rank, rank)
        xchg = DeviceTrajectoryExchange(eng, torch.device("cuda", rank))
        records, counts = xchg.gather()
        assert records.is_cuda and records.shape[1] == 216 and len(counts) == world
        got = records.cpu().numpy().reshape(-1).view(tb._lib.EXAMPLE_DTYPE)
        want = want_parts = []          # noqa: F841
        for r in range(world):                       # identical engines, drained through the host path
            twin = play(tb, 48 + 16 * r, sims, 100 + r, rank)
            want.append(twin.selfplay_drain())
            twin.close()
        assert counts == [len(w) for w in want_parts]
        at = 0
        for part in want_parts:                      # rank order; inside a rank, games ending on the same ply may swap
            mine = got[at:at + len(part)]
            at += len(part)
            a, b = np.lexsort((mine["ply"], mine["env"])), np.lexsort((part["ply"], part["env"]))
            for name in tb._lib.EXAMPLE_DTYPE.names:
                assert np.array_equal(mine[name][a], part[name][b]), name
        # learner step from the device replay: both ranks draw the same global batch, take their shard, all-reduce
        replay = DeviceReplay(eng, records)
        torch.manual_seed(0)
        learner = Learner(TrunkNet(16, 1, 64, dropout=0.0), lr=1e-3, device=f"cuda:{rank}")
        gen = torch.Generator().manual_seed(5)
        idx = torch.randint(0, replay.n_examples, (256,), generator=gen)
        obs, pi, z = replay.examples(idx[rank::world].cuda())
        ho, hp, hz = records_to_arrays(got, augment=True)
        k = idx[rank::world].numpy()
        assert np.array_equal(obs.cpu().numpy(), ho[k]) and np.array_equal(pi.cpu().numpy(), hp[k])
        losses = learner.training_step(obs, pi, z, global_batch=256)
        assert all(np.isfinite(losses))
        flat = torch.cat([p.detach().reshape(-1) for p in learner.net.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])         # same weights on every rank after the step
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_device_records_all_gather_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(2, free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
