#!/usr/bin/env python
"""Learner step throughput: `Learner.training_step` on synthetic examples, trainer-default net.

    python tools/learner_bench.py [--batch 8192] [--steps 30]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/learner_bench.py   # data parallel

`--batch` is the GLOBAL batch (nn_wrapper.py's args.batch_size); every rank takes batch / N rows.  Reports examples/s
over all ranks (CUDA-event time, max over ranks) and the share of the step spent in the gradient all-reduce."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from tacult_b200.learner import Learner  # noqa: E402
from tacult_b200.network import TrunkNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    net = TrunkNet(32, 3, 256, dropout=0.2)
    learner = Learner(net, lr=1e-3, device=f"cuda:{local}")
    learner.broadcast(0)
    learner.net.train()
    rows = a.batch // world
    rng = np.random.RandomState(rank)
    obs = torch.from_numpy((rng.random_sample((rows, 4, 9, 9)) < 0.3).astype(np.float32)).cuda()
    pi = torch.from_numpy(rng.dirichlet(np.ones(81), rows).astype(np.float32)).cuda()
    z = torch.from_numpy(rng.choice([-1.0, 1.0, 1e-4], rows).astype(np.float32)).cuda()
    for _ in range(a.warmup):
        learner.training_step(obs, pi, z, global_batch=a.batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        learner.training_step(obs, pi, z, global_batch=a.batch)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # all-reduce alone, same buffer
    ar_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            dist.all_reduce(learner._flat_grad)
        torch.cuda.synchronize()
        ar_ms = 1e3 * (time.perf_counter() - t0) / 20
    if rank == 0:
        per_step = float(ms.item()) / a.steps
        print(json.dumps({"workload": f"learner step, trainer-default net, global batch {a.batch}, Adam + clip-norm 1.0",
                          "n_gpus": world, "ms_per_step": per_step, "examples_per_sec": a.batch / (per_step * 1e-3),
                          "params": int(learner._flat_grad.numel()), "grad_all_reduce_ms": ar_ms,
                          "grad_all_reduce_share": ar_ms / per_step if per_step else None,
                          "note": "autograd graph is torch (library); one flat-buffer NCCL all-reduce per step"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
