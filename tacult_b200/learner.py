"""Data-parallel learner step and checkpoint interchange.

`Learner.training_step` is `NNetWrapper.training_step` (/root/reference/src/tacult/base/nn_wrapper.py:45-56):
    pi_loss = -sum(pi_target * log(pi + 1e-8)) / B,  v_loss = mean((z - v)^2),  loss = pi_loss + v_loss,
    backward, clip_grad_norm_(1.0), Adam(lr) step
and `Learner.train` its fixed-step epoch loop (nn_wrapper.py:58-122).  With a process group every rank holds a
shard of the global batch: the local loss is divided by the GLOBAL batch size and the gradients are summed with
ONE all-reduce over a flat buffer that all `.grad` tensors alias (NCCL on GPUs, gloo in the CPU tests), so the
step equals the single-process step on the concatenated batch (BatchNorm batch statistics stay per rank, as in
any data-parallel trainer; with one rank the arithmetic is the reference's, op for op).

`publish(engine)` hands the new weights to the self-play engine (`tc_load_weights` re-folds BatchNorm), which is
what `Coach` does through `load_checkpoint` after a training round (coach.py:217-219,256).

Checkpoints use the reference's file format — `torch.save({'state_dict', 'optimizer'})`, nn_wrapper.py:160-163 —
so `.pt` files move both ways.  The autograd graph is torch (library code); nothing here is on the self-play path.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .network import TrunkNet
from .weights import _strip


class Learner:
    def __init__(self, net, lr, device="cpu", group=None, distributed=None):
        self.device = torch.device(device)
        self.net = net.to(self.device)
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else bool(distributed)
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=lr)          # nn_wrapper.py:33
        # every .grad is a view into one flat buffer: one collective per step, no packing copies
        params = [p for p in self.net.parameters() if p.requires_grad]
        self._flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=self.device)
        offset = 0
        for p in params:
            p.grad = self._flat_grad[offset:offset + p.numel()].view_as(p)
            offset += p.numel()
        self.steps = 0

    # -- one optimisation step ------------------------------------------------------------------
    def training_step(self, boards, pis, vs, global_batch=None):
        """boards f32[b,4,9,9], pis f32[b,81], vs f32[b]: this rank's shard.  Returns (pi_loss, v_loss) of the
        GLOBAL batch (summed over ranks)."""
        boards, pis, vs = boards.to(self.device), pis.to(self.device), vs.to(self.device)
        local = boards.shape[0]
        if global_batch is None:
            global_batch = local
            if self.distributed:
                n = torch.tensor([local], dtype=torch.int64, device=self.device)
                dist.all_reduce(n, group=self.group)
                global_batch = int(n.item())
        self._flat_grad.zero_()
        out_pi, out_v = self.net(boards)
        pi_loss = -torch.sum(pis * torch.log(out_pi + 1e-8)) / global_batch
        if self.world == 1:
            v_loss = torch.mean((vs - out_v.squeeze()) ** 2)                     # the reference's expression, op for op
        else:
            v_loss = torch.sum((vs - out_v.reshape(-1)) ** 2) / global_batch
        (pi_loss + v_loss).backward()
        losses = torch.stack([pi_loss.detach(), v_loss.detach()])
        if self.distributed and self.world > 1:
            dist.all_reduce(self._flat_grad, group=self.group)                   # sum of per-shard gradients
            dist.all_reduce(losses, group=self.group)
        torch.nn.utils.clip_grad_norm_(self.net.parameters(), 1.0)
        self.optimizer.step()
        self.steps += 1
        return float(losses[0]), float(losses[1])

    # -- the epoch loop ----------------------------------------------------------------------------
    def train(self, obs, pi, z, epochs, steps_per_epoch, batch_size, shuffle=True, seed=0):
        """obs f32[n,4,9,9], pi f32[n,81], z f32[n] (numpy or tensors; the same arrays on every rank).
        `batch_size` is the global batch; every rank takes rows rank::world of each batch.  Batches walk a
        permutation of the examples that restarts when exhausted, like the reference's persistent DataLoader
        iterator (nn_wrapper.py:72-99).  Returns [(mean pi_loss, mean v_loss)] per epoch."""
        obs, pi, z = (torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in (obs, pi, z))
        n = obs.shape[0]
        gen = torch.Generator().manual_seed(int(seed))          # same permutation on every rank
        order, cursor = None, 0
        history = []
        self.net.train()
        for _ in range(int(epochs)):
            tot_pi = tot_v = 0.0
            for _ in range(int(steps_per_epoch)):
                if order is None or cursor >= n:
                    order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
                    cursor = 0
                idx = order[cursor:cursor + batch_size]
                cursor += batch_size
                if len(idx) < 2 * self.world and n >= 2 * self.world:
                    # ragged tail too short to give every rank the two rows train-mode BatchNorm needs (a rank that
                    # raised there would leave the others hanging in the all-reduce): top it up from the next permutation
                    short = len(idx)
                    order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
                    cursor = min(n, batch_size - short)
                    idx = torch.cat([idx, order[:cursor]])
                mine = idx[self.rank::self.world]
                a, b = self.training_step(obs[mine], pi[mine], z[mine], global_batch=len(idx))
                tot_pi += a
                tot_v += b
            history.append((tot_pi / steps_per_epoch, tot_v / steps_per_epoch))
        return history

    def train_from_replay(self, replay, epochs, steps_per_epoch, batch_size, shuffle=True, seed=0):
        """The same epoch loop fed from a `distributed.DeviceReplay` (trajectory records resident in HBM): every step
        draws the global batch of example ids (record x 8 + symmetry image) from the shared permutation, takes this
        rank's rows and lets `tc_examples_from_records_dev` build exactly those (obs, pi, z) on the device — the
        eight-fold augmented training set (coach.py:136-138) is never materialised."""
        n = replay.n_examples
        gen = torch.Generator().manual_seed(int(seed))          # same permutation on every rank
        order, cursor = None, 0
        history = []
        self.net.train()
        for _ in range(int(epochs)):
            tot_pi = tot_v = 0.0
            for _ in range(int(steps_per_epoch)):
                if order is None or cursor >= n:
                    order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
                    cursor = 0
                idx = order[cursor:cursor + batch_size]
                cursor += batch_size
                if len(idx) < 2 * self.world and n >= 2 * self.world:
                    short = len(idx)
                    order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
                    cursor = min(n, batch_size - short)
                    idx = torch.cat([idx, order[:cursor]])
                obs, pi, z = replay.examples(idx[self.rank::self.world])
                a, b = self.training_step(obs, pi, z, global_batch=len(idx))
                tot_pi += a
                tot_v += b
            history.append((tot_pi / steps_per_epoch, tot_v / steps_per_epoch))
        return history

    # -- weights out / in --------------------------------------------------------------------------
    def state_dict_numpy(self):
        return {k: v.detach().cpu().numpy() for k, v in self.net.state_dict().items()}

    def publish(self, engine):
        """New weights -> self-play engine (BatchNorm folded again inside tc_load_weights)."""
        engine.load_weights(self.state_dict_numpy())

    def broadcast(self, src=0):
        """Rank `src`'s parameters and BatchNorm buffers to every rank (accepted-model broadcast)."""
        if self.distributed and self.world > 1:
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def save_checkpoint(self, folder="checkpoint", filename="checkpoint.pt"):
        os.makedirs(folder, exist_ok=True)
        torch.save({"state_dict": self.net.state_dict(), "optimizer": self.optimizer.state_dict()},
                   os.path.join(folder, filename))

    def load_checkpoint(self, folder="checkpoint", filename="checkpoint.pt"):
        path = os.path.join(folder, filename)
        if not os.path.exists(path):
            raise FileNotFoundError(f"No model found at {path}")                 # nn_wrapper.py:169-171
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.net.load_state_dict(_strip(ckpt["state_dict"]))      # torch.compile'd reference nets prefix "_orig_mod."
        if "optimizer" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer"])


def load_reference_checkpoint(path, map_location="cpu"):
    """A reference `.pt` (nn_wrapper.py:160-163) -> (state_dict of numpy arrays, optimizer state or None).
    Feed the state_dict to `Engine.load_weights` or `TrunkNet.from_state_dict`."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = {k: v.detach().cpu().numpy() for k, v in _strip(ckpt["state_dict"]).items()}
    return sd, ckpt.get("optimizer")


def learner_from_checkpoint(path, lr, device="cpu", dropout=0.3, **kw):
    sd, opt = load_reference_checkpoint(path)
    learner = Learner(TrunkNet.from_state_dict(sd, dropout=dropout), lr, device=device, **kw)
    if opt is not None:
        learner.optimizer.load_state_dict(opt)
    return learner
