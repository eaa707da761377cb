"""Multi-GPU plumbing: one process per GPU, trees sharded by env index, no collective on the search path.

The reference is single-process; what its `Coach.generateTrainExamples`
(/root/reference/src/tacult/base/coach.py:160-193) returns — the iteration's training examples — is here the
concatenation over ranks of every rank's finished trajectories, exchanged once per iteration:
counts are all-gathered first, then one padded all-gather of fixed-size `tc_example` records (216 bytes each;
the 8-fold symmetry augmentation of coach.py:136-138 is applied after the gather, so it is never sent).
Works over NCCL (records stay on the device: tc_selfplay_drain_dev) and over gloo (CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import EXAMPLE_DTYPE

RECORD_BYTES = EXAMPLE_DTYPE.itemsize


def shard_envs(total_envs, world_size, rank):
    """Contiguous block of env indices owned by `rank`: [start, stop)."""
    base, extra = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_records(local, group=None, device=None):
    """local: uint8 tensor [n_local * 216] (CPU for gloo, CUDA for nccl) or a numpy record array.
    Returns the records of all ranks, in rank order, as a numpy array of EXAMPLE_DTYPE."""
    if isinstance(local, np.ndarray):
        assert local.dtype == EXAMPLE_DTYPE
        local = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1).copy())
    if device is not None:
        local = local.to(device)
    world = dist.get_world_size(group)
    n_local = local.numel() // RECORD_BYTES
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    mine = torch.tensor([n_local], dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_host = counts.cpu().tolist()
    n_max = max(counts_host) if counts_host else 0
    if n_max == 0:
        return np.zeros(0, dtype=EXAMPLE_DTYPE)
    padded = torch.zeros(n_max * RECORD_BYTES, dtype=torch.uint8, device=local.device)
    padded[: local.numel()] = local
    gathered = torch.empty(world * n_max * RECORD_BYTES, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    host = gathered.cpu().numpy().reshape(world, n_max * RECORD_BYTES)
    parts = [host[r, : counts_host[r] * RECORD_BYTES] for r in range(world)]
    return np.concatenate(parts).view(EXAMPLE_DTYPE).copy()


def broadcast_weights(module, src=0, group=None):
    """Accepted weights go from the learner rank to every actor rank (coach.py:246-256 accept step)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def all_reduce_gradients(module, group=None):
    """Data-parallel learner step: average gradients over ranks before clip + Adam (nn_wrapper.py:45-55)."""
    world = dist.get_world_size(group)
    for p in module.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
            p.grad.div_(world)
