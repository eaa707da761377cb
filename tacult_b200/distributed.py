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


class DeviceTrajectoryExchange:
    """The iteration's one exchange step with the records resident in HBM from the self-play kernel to the learner:

        tc_selfplay_drain_dev  ->  persistent CUDA send buffer  ->  all_gather_into_tensor (NCCL over NVLink)
                               ->  compacted record tensor on the device  ->  tc_examples_from_records_dev

    Only the per-rank record COUNTS visit the host (one 8-byte-per-rank read to size the collective).  The send
    buffer is allocated once (E x 256 records, the engine's own finished-record capacity) and reused every
    iteration; ranks send their first max(count) slots, so no padding copy is made."""

    def __init__(self, engine, device, group=None, capacity=None):
        self.engine = engine
        self.device = torch.device(device)
        self.group = group
        self.capacity = int(capacity or engine.num_envs * 256)
        self.send = torch.empty(self.capacity * RECORD_BYTES, dtype=torch.uint8, device=self.device)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        self._mine = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._recv = self._compact = None

    def warm_up(self, records_per_rank=None):
        """Allocates the receive buffer for `records_per_rank` records from every rank (default: one game of 81 plies per
        env) and runs both collectives once, so that the first real exchange of an iteration pays neither cudaMalloc nor
        NCCL's lazy channel set-up."""
        n = int(records_per_rank or min(self.capacity, self.engine.num_envs * 81))
        if self.world == 1:
            return
        need = self.world * n * RECORD_BYTES
        if self._recv is None or self._recv.numel() < need:
            self._recv = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._compact = torch.empty(need, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(self._counts, self._mine, group=self.group)
        dist.all_gather_into_tensor(self._recv[:need], self.send[: n * RECORD_BYTES], group=self.group)
        torch.cuda.synchronize(self.device)

    def drain(self):
        """This rank's finished-game records into the send buffer; returns the count."""
        return self.engine.selfplay_drain_dev(self.send.data_ptr(), self.capacity)

    def gather(self, n_local=None):
        """Returns (records uint8[n_total, 216] on the device, list of per-rank counts)."""
        if n_local is None:
            n_local = self.drain()
        if self.world == 1:
            return self.send[: n_local * RECORD_BYTES].view(n_local, RECORD_BYTES), [n_local]
        self._mine.fill_(n_local)
        dist.all_gather_into_tensor(self._counts, self._mine, group=self.group)
        counts = self._counts.tolist()
        n_max = max(counts)
        if n_max == 0:
            return self.send[:0].view(0, RECORD_BYTES), counts
        need = self.world * n_max * RECORD_BYTES
        if self._recv is None or self._recv.numel() < need:
            self._recv = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._compact = torch.empty(need, dtype=torch.uint8, device=self.device)
        recv = self._recv[:need]
        dist.all_gather_into_tensor(recv, self.send[: n_max * RECORD_BYTES], group=self.group)
        rows = recv.view(self.world, n_max, RECORD_BYTES)
        if all(c == n_max for c in counts):
            return rows.view(self.world * n_max, RECORD_BYTES), counts
        # ragged: close the gaps with device-to-device copies into the persistent compact buffer (no allocation)
        out = self._compact[: sum(counts) * RECORD_BYTES].view(sum(counts), RECORD_BYTES)
        at = 0
        for r in range(self.world):
            out[at:at + counts[r]].copy_(rows[r, : counts[r]])
            at += counts[r]
        return out, counts


class DeviceReplay:
    """Replay buffer of raw trajectory records in HBM (216 bytes per env-ply).  The reference keeps eight float
    images of every position in a host deque (coach.py:136-138,38: ~13 KB per env-ply); here the images are
    produced on demand by the example kernel for exactly the rows a learner step asks for."""

    def __init__(self, engine, records_dev):
        self.engine = engine
        self.records = records_dev.contiguous()
        self.n_records = int(self.records.shape[0])
        self.n_examples = 8 * self.n_records

    def examples(self, index=None):
        """index: int32/int64 tensor of example ids (record * 8 + image) on the records' device, or None for all of
        them in order.  Returns (obs f32[m,4,9,9], pi f32[m,81], z f32[m]) on the device."""
        dev = self.records.device
        if index is None:
            m, idx_ptr = self.n_examples, 0
        else:
            index = index.to(device=dev, dtype=torch.int32).contiguous()
            m, idx_ptr = int(index.numel()), index.data_ptr()
        obs = torch.empty((m, 4, 9, 9), dtype=torch.float32, device=dev)
        pi = torch.empty((m, 81), dtype=torch.float32, device=dev)
        z = torch.empty((m,), dtype=torch.float32, device=dev)
        if m:
            # the kernel runs on the engine's stream: order it after the producer of `records` / `index` and make the
            # consumer (torch's current stream) wait for it
            torch.cuda.current_stream(dev).synchronize()
            self.engine.examples_from_records_dev(m, self.records.data_ptr(), self.n_records, idx_ptr,
                                                  obs.data_ptr(), pi.data_ptr(), z.data_ptr())
            self.engine.sync()
        return obs, pi, z


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
