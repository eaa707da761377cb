#!/usr/bin/env python
"""Leaf-evaluation throughput sweep (BASELINE.json configs[4]: deep trunk 20 blocks x 256 ch, batch 256..65536).

    python tools/eval_sweep.py --net deep --evaluator bf16 --batches 256,1024,4096,16384,65536
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_sweep.py --net deep ...
    python tools/eval_sweep.py --net deep --in-planes 32        # the 8-frame history variant (C_in = 32)

Inputs (packed positions, or dense history planes for --in-planes 32) and outputs are resident in HBM; one JSON line per
batch size with evals/s and the achieved fraction of the measured bf16 tensor peak (algorithmic FLOPs of SURVEY.md §8d /
CUDA-event time).  Under torchrun the weights are replicated, `batch` is the GLOBAL batch sharded evenly over the ranks
(no collective on the path), the time is the max over ranks and evals/s the whole-box figure."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import tacult_b200 as tb  # noqa: E402
from tacult_b200.weights import flops_per_eval, random_state_dict  # noqa: E402

NETS = {"deep": (256, 20, 256), "trainer": (32, 3, 256), "class": (16, 4, 256), "wide64": (64, 2, 128)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="deep", choices=sorted(NETS))
    ap.add_argument("--evaluator", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batches", default="256,1024,4096,16384,65536")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true", help="per-kernel-class CUDA-event times")
    ap.add_argument("--in-planes", type=int, default=4)
    a = ap.parse_args()
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C, B, E = NETS[a.net]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    flops = flops_per_eval(C, B, E, a.in_planes)
    ev = tb._lib.EVAL_NET_BF16 if a.evaluator == "bf16" else tb._lib.EVAL_NET_F32
    sd = random_state_dict(0, C, B, E, a.in_planes)
    stream = torch.cuda.Stream()          # non-default stream shared by the engine and the timing events
    torch.cuda.set_stream(stream)
    for global_batch in [int(x) for x in a.batches.split(",")]:
        batch = max(1, global_batch // world)
        eng = tb.Engine(batch, 2, device=local)
        eng.set_stream(stream.cuda_stream)
        eng.load_weights(sd)
        # synthetic positions: random playouts on the device give real mid-game boards
        res = tb.rules_playouts(1, 0, min(batch, 65536), max_plies=30)
        states = np.array([list(r.final_state) for r in res], dtype=np.uint32)
        states = np.resize(states, (batch, 8))
        d_states = torch.from_numpy(states.view(np.int32)).cuda()
        d_obs = None
        if a.in_planes != 4:        # dense planes: frame f of position i = position (i + f) of the batch (synthetic history)
            frames = a.in_planes // 4
            hist = np.stack([np.roll(states, -f, axis=0) for f in range(frames)], axis=1)
            d_obs = torch.from_numpy(tb.rules_encode_history(hist)).cuda()
        d_pi = torch.empty((batch, 81), dtype=torch.float32, device="cuda")
        d_v = torch.empty((batch,), dtype=torch.float32, device="cuda")

        def run():
            if d_obs is None:
                tb._lib.check(eng.lib.tc_forward_states_dev(eng.handle, ev, batch, d_states.data_ptr(), d_pi.data_ptr(),
                                                           d_v.data_ptr()))
            else:
                tb._lib.check(eng.lib.tc_forward_obs_dev(eng.handle, ev, batch, d_obs.data_ptr(), d_pi.data_ptr(),
                                                        d_v.data_ptr()))
        for _ in range(a.warmup):
            run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        tf_gpu = flops * batch / (ms * 1e-3) / 1e12
        line = {"net": a.net, "channels": C, "blocks": B, "in_planes": a.in_planes, "evaluator": a.evaluator,
                "n_gpus": world, "batch": batch * world, "batch_per_gpu": batch, "ms": ms,
                "evals_per_sec": batch * world / (ms * 1e-3), "tflops_per_gpu": tf_gpu,
                "frac_of_bf16_burst_peak": tf_gpu / peaks["bf16_tflops"],
                "frac_of_bf16_sustained_peak": tf_gpu / peaks["bf16_tflops_sustained"],
                "pi_checksum": float(d_pi.sum().item())}
        if a.profile:
            eng.profile_enable(True)
            for _ in range(3):
                run()
            line["kernels_ms_per_iter"] = {k: v[0] / 3 for k, v in eng.profile_read().items() if v[1]}
            eng.profile_enable(False)
        if rank == 0:
            print(json.dumps(line), flush=True)
        eng.close()
        del d_states, d_pi, d_v, d_obs
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
