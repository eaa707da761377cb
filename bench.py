#!/usr/bin/env python
"""Benchmark of the self-play hot path (BASELINE.json metric: MCTS sims/sec, self-play, whole box).

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

A "step" is one lock-step ply of every tree: `sims` simulations per live tree (select -> leaf evaluation
-> expand/backup), move choice from the visit counts, root advance.  Workload at N=1: BASELINE.json
configs[2] — 4096 trees x 200 sims/move, trainer-default ResNet (32 ch x 3 blocks, emb 256), random-init
weights, synthetic self-play.  N>1: every rank runs its own 4096 trees (weak scaling, no data-path
collective; finished trajectories are all-gathered outside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET = dict(channels=32, num_residual_blocks=3, embedding_size=256)
CPUCT = 2.4
TEMP_THRESHOLD = 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--evaluator", default="auto", choices=["auto", "f32", "bf16"])
    ap.add_argument("--stagger-plies", type=int, default=150)
    ap.add_argument("--profile-plies", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-envs", type=int, default=16, help="trees in the host-CPU sample (reference arm)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        ok = [r for r in self.rows if len(r) == 6 and r[0].isdigit()]
        if not ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in ok:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([int(r[0]) for r in ok])), "sm_max_mhz": float(ok[0][1]),
                "reasons": sorted(reasons), "samples": len(ok)}


# ------------------------------------------------------------------------------------------------
# host-CPU arm: the reference algorithm (oracle port; /root/reference is Python and cannot travel)
# ------------------------------------------------------------------------------------------------
def cpu_selfplay_rate(envs, sims, seconds=None, plies=None, warmup_plies=0, seed=0):
    """Times the reference's CPU path — lock-step VectorizedMCTS + rules engine + torch-CPU ResNet forward —
    as restated in oracle/ (mcts_oracle.py, game_oracle.py, resnet_oracle.py), on `envs` trees x `sims` sims.
    Returns (sims_per_second, plies_timed, seconds_timed, per_ply_ms list)."""
    import logging
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from game_oracle import OracleGame
    from mcts_oracle import OracleVectorizedMCTS
    import resnet_oracle
    logging.disable(logging.ERROR)
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(seed)

    class Args(dict):
        __getattr__ = dict.__getitem__
    args = Args(numEnvs=envs, numMCTSSims=sims, cpuct=CPUCT)
    game = OracleGame()
    net = resnet_oracle.OracleNet(resnet_oracle.make_state_dict(0, **NET))
    mcts = OracleVectorizedMCTS(game, net, args)
    boards = [game.getInitBoard() for _ in range(envs)]
    players = np.ones(envs, dtype=int)
    step_no = np.zeros(envs, dtype=int)

    def one_ply():
        nonlocal mcts
        live = 0
        step_no[:] += 1
        canon = [game.getCanonicalForm(boards[i], players[i]) for i in range(envs)]
        probs = mcts.getActionProbs(canon, temps=(step_no < TEMP_THRESHOLD))
        for i in range(envs):
            if game.getGameEnded(boards[i], players[i]) != 0:      # restart finished games (keeps the sample busy)
                boards[i], players[i], step_no[i] = game.getInitBoard(), 1, 0
                mcts.reset(i)
                continue
            live += 1
            a = np.random.choice(81, p=probs[i])
            boards[i], players[i] = game.getNextState(boards[i], int(players[i]), a)
        return live

    for _ in range(warmup_plies):
        one_ply()
    t0 = time.perf_counter()
    done_sims, n, per_ply = 0, 0, []
    while True:
        t1 = time.perf_counter()
        done_sims += one_ply() * sims
        per_ply.append((time.perf_counter() - t1) * 1e3)
        n += 1
        if plies is not None and n >= plies:
            break
        if seconds is not None and time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return done_sims / dt, n, dt, per_ply


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    rate, n, dt, per_ply = cpu_selfplay_rate(a.ref_envs, a.sims, plies=a.steps, warmup_plies=a.warmup)
    cores = torch.get_num_threads()
    sample = (f"{a.ref_envs} of the {a.envs} trees x {a.sims} sims/move, {a.steps} plies after {a.warmup} warm-up, "
              f"oracle port of VectorizedMCTS+rules+torch-CPU ResNet")
    line = {
        "impl": "reference", "metric": "mcts_sims_per_sec_selfplay", "value": rate, "unit": "sims/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(per_ply)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, "cpu"),
        "cpu_baseline": {"value": rate, "unit": "sims/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": rate, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a, evaluator):
    return {"workload": f"self-play {a.envs} trees x {a.sims} sims/move per GPU, trainer-default ResNet "
                        f"(32ch x 3 blocks, emb 256), random init",
            "trees_per_gpu": a.envs, "sims_per_move": a.sims, "cpuct": CPUCT, "temp_threshold": TEMP_THRESHOLD,
            "evaluator": evaluator, "tree_arithmetic": "f64",
            "stagger": f"{a.stagger_plies} plies @2 sims with auto-reset before warm-up",
            "l2": "working set (tree store + activations) exceeds the 126 MB L2; no explicit flush",
            "parallelism": f"trees sharded by index, {a.gpus} x {a.envs}"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import tacult_b200 as tb
    from tacult_b200.weights import flops_per_eval, random_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_dev = local
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_scalar(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    E, sims = a.envs, a.sims
    eng = tb.Engine(E, sims * 82 + 64, device=local)
    # a real (non-default) torch stream: the engine launches on it, and the timing events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    eng.load_weights(random_state_dict(0, **NET))
    evaluator = a.evaluator
    if evaluator == "auto":
        try:
            eng.set_evaluator(tb._lib.EVAL_NET_BF16)
            eng.selfplay_begin(2, CPUCT, TEMP_THRESHOLD, seed=1, auto_reset=True)
            evaluator = "bf16"
        except tb._lib.TacultError:
            evaluator = "f32"
    eng.set_evaluator(tb._lib.EVAL_NET_BF16 if evaluator == "bf16" else tb._lib.EVAL_NET_F32)

    # ---- device-resident self-play: the headline `value` --------------------------------------
    seed = 1000 + rank
    eng.selfplay_begin(2, CPUCT, TEMP_THRESHOLD, seed=seed, auto_reset=True, discard_records=True)
    eng.selfplay_step(a.stagger_plies)                       # spread the envs over all game phases
    eng.selfplay_configure(sims, CPUCT, TEMP_THRESHOLD, seed=seed, auto_reset=True)   # records kept from here on
    eng.selfplay_step(a.warmup)
    eng.selfplay_sync()
    eng.selfplay_drain()
    s0, k0 = eng.selfplay_stats(), eng.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        eng.selfplay_step(a.steps)
        ev1.record(stream)
        torch.cuda.synchronize()
    barrier()
    eng.selfplay_sync()
    ms = reduce_scalar(ev0.elapsed_time(ev1), dist.ReduceOp.MAX if world > 1 else None)
    s1, k1 = eng.selfplay_stats(), eng.stats()
    sims_done = reduce_scalar(float(s1["sims"] - s0["sims"]), dist.ReduceOp.SUM if world > 1 else None)
    leaves = k1["nn_leaves"] - k0["nn_leaves"]
    launches = k1["kernel_launches"] - k0["kernel_launches"]
    value = sims_done / (ms * 1e-3)

    # ---- per-kernel device time (CUDA events around every launch) for the roofline -------------
    eng.selfplay_drain()
    kp0 = eng.stats()
    eng.profile_enable(True)
    eng.selfplay_step(a.profile_plies)
    eng.selfplay_sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    kp1 = eng.stats()
    roof, kernels = None, {}
    if a.profile_plies > 0:
        roof, kernels = roofline(prof, kp0, kp1, a, evaluator, flops_per_eval(NET["channels"], NET["num_residual_blocks"],
                                                                              NET["embedding_size"]))

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    e2e = None
    if not a.skip_e2e:
        e2e = run_e2e(a, eng, tb, torch, rank, barrier, reduce_scalar, dist, world)

    # ---- N>1: the one exchange step — finished trajectories all-gathered over NCCL (outside the timed region)
    gathered = None
    if world > 1:
        from tacult_b200.distributed import all_gather_records
        eng.selfplay_configure(sims, CPUCT, TEMP_THRESHOLD, seed=seed, auto_reset=True)
        eng.selfplay_drain()
        eng.selfplay_step(2)
        eng.selfplay_sync()
        local_recs = eng.selfplay_drain()
        t0 = time.perf_counter()
        allrec = all_gather_records(local_recs, device=torch.device("cuda", local_dev))
        torch.cuda.synchronize()
        gathered = {"records_local": int(len(local_recs)), "records_all_ranks": int(len(allrec)),
                    "ms": 1e3 * (time.perf_counter() - t0), "bytes_per_record": 216}

    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        rate, n, dt, _ = cpu_selfplay_rate(64, 25, seconds=a.cpu_seconds)
        cpu = {"value": rate, "unit": "sims/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"BASELINE config C1 shape: 64 trees x 25 sims/move, {n} plies in {dt:.1f} s, oracle port of "
                         f"VectorizedMCTS+rules+torch-CPU ResNet (python search loop is single-threaded)",
               "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": "mcts_sims_per_sec_selfplay", "value": value, "unit": "sims/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": evaluator, "data": "synthetic",
            "config": workload_config(a, evaluator),
            "leaf_evals_per_sec": leaves * world / (ms * 1e-3),
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roof, "kernels": kernels,
            "e2e": e2e, "cpu_baseline": cpu, "trajectory_all_gather": gathered,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline(prof, k0, k1, a, evaluator, flops_eval):
    peaks = measured_peaks()
    total_ms = sum(v[0] for v in prof.values()) or 1e-9
    d = {k: k1[k] - k0[k] for k in k0}
    kernels = {}
    for name, (ms, n) in prof.items():
        if n:
            kernels[name] = {"ms": ms, "launches": n, "share": ms / total_ms, "avg_us": 1e3 * ms / n}
    sims_steps = max(d["sims"], 1)
    # algorithmic bytes of the tree side per simulation: select reads P,Q,N per legal edge + header,
    # backup RMW per level, expansion writes header + edges, reads pi/v, writes the 32-byte leaf position
    # (layout in DESIGN.md §3): select reads the 32-byte edge records of every level, writes the path, and at a new
    # leaf reads the parent header, probes 128 B of the hash table, writes header + 16 B per edge + the 32-byte leaf
    sel_bytes = 32.0 * d["sum_legal"] + 8.0 * d["sum_depth"] + (32.0 + 128.0 + 32.0 + 32.0) * d["nn_leaves"] \
        + 16.0 * d["sum_leaf_legal"]
    # expand/backup: header + pi/v of the leaf, 8 B prior per edge, 24 B read-modify-write per path level
    exp_bytes = (32.0 + 328.0) * d["nn_leaves"] + 8.0 * d["sum_leaf_legal"] + 32.0 * d["sum_depth"]
    # trunk = conv_in + the residual-block convolutions (one fused launch for this net): SURVEY.md §8d trunk term
    trunk_flops_per_leaf = 2.0 * 81 * 9 * NET["channels"] * (4 + 2 * NET["num_residual_blocks"] * NET["channels"])
    conv_flops = trunk_flops_per_leaf * d["nn_leaves"]
    # the backup of simulation k and the descent of simulation k+1 share one launch (class "select"); the separate
    # expand_backup launch only closes the last simulation of a move, so the tree-side bytes are attributed together
    tree_ms = kernels.get("select", {}).get("ms", 0.0) + kernels.get("expand_backup", {}).get("ms", 0.0)
    if "select" in kernels:
        kernels["select"].update(bound="hbm", achieved=(sel_bytes + exp_bytes) / (tree_ms * 1e-3) / 1e9, unit="GB/s",
                                 note="descent + fused backup/expansion of the previous simulation")
    if "expand_backup" in kernels:
        kernels["expand_backup"].update(bound="hbm", note="last simulation of a move only; bytes counted under select")
    if "nn_conv" in kernels:
        kernels["nn_conv"].update(bound="tensor", achieved=conv_flops / (kernels["nn_conv"]["ms"] * 1e-3) / 1e12,
                                  unit="TFLOP/s")
    nn_ms = sum(kernels[k]["ms"] for k in kernels if k.startswith("nn_"))
    if nn_ms:
        kernels["nn_all"] = {"ms": nn_ms, "share": nn_ms / total_ms, "bound": "tensor", "unit": "TFLOP/s",
                             "achieved": flops_eval * d["nn_leaves"] / (nn_ms * 1e-3) / 1e12}
    top = max((k for k in kernels if k != "nn_all"), key=lambda k: kernels[k]["ms"])
    tk = kernels[top]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top)
    if tk.get("bound") == "hbm":
        peak = peaks["hbm_gbs"]
    else:
        peak = peaks["bf16_tflops_sustained"]
    roof = {"kernel": top, "bound": tk.get("bound", "tensor"), "achieved": tk.get("achieved"), "peak": peak,
            "unit": tk.get("unit", "TFLOP/s"), "frac": (tk.get("achieved") or 0.0) / peak, "traffic": traffic,
            "peak_source": peaks["source"], "avg_launch_us": tk["avg_us"], "share_of_step": tk["share"],
            "rows_per_launch": d["nn_leaves"] / max(kernels.get("nn_conv", {}).get("launches", 0), 1)
            if top.startswith("nn_") else None,
            "algorithmic_flops_per_leaf": trunk_flops_per_leaf if top == "nn_conv" else None}
    if top == "nn_conv":
        # measured with tools/probes/umma_rate_probe.cu: a 128 x N x 16 UMMA costs max((4096 + 32 N) / 128, N / 2) cycles, i.e.
        # at N = 32 channels the tensor pipe is fed by shared-memory operand reads at 128 B/clk, not by its math rate
        roof["note"] = ("operand-fetch bound at 32 channels: 40 cycles per UMMA for 16 cycles of math; the kernel runs at "
                        "about 82 % of that shared-memory bound (DESIGN.md section 4)")
    return roof, kernels


def run_e2e(a, eng, tb, torch, rank, barrier, reduce_scalar, dist, world):
    """Same metric through the C ABI with HOST buffers: every ply copies the root positions host->device,
    searches, reads the visit counts device->host, picks moves on the host and advances the positions with
    tc_rules_step (host buffers again).  Inputs live in pinned host memory."""
    E, sims = a.envs, a.sims
    rng = np.random.RandomState(7 + rank)
    states = torch.zeros((E, 8), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    # staggered synthetic positions: random legal moves to a random depth below 50 plies
    depth = rng.randint(0, 50, size=E)
    ply_no = np.zeros(E, dtype=np.int64)
    for d in range(50):
        todo = np.flatnonzero(depth > d)
        if todo.size == 0:
            break
        mask = tb.rules_legal_mask(states[todo])
        acts = np.array([rng.choice(np.flatnonzero(m)) if m.any() else -1 for m in mask], dtype=np.int8)
        ok = acts >= 0
        nxt, st = tb.rules_step(states[todo][ok], acts[ok])
        keep = st == 0                          # do not step into finished games
        idx = todo[ok][keep]
        states[idx] = nxt[keep]
        ply_no[idx] += 1
    eng.reset(-1)
    eng.reset_stats()
    h2d = E * 32 + E * 32 + E          # roots + positions and actions handed to tc_rules_step
    d2h = E * 81 * 4 + E * 32 + E      # visit counts + successor positions and their status

    def one_ply():
        eng.set_roots_packed(states)
        eng.search(sims, CPUCT)
        counts = eng.root_counts()
        tau_one = (ply_no + 1) < TEMP_THRESHOLD
        cum = np.cumsum(counts, axis=1, dtype=np.int64)
        draw = (rng.random_sample(E) * cum[:, -1]).astype(np.int64)
        sampled = (cum > draw[:, None]).argmax(1)               # a ~ counts / sum (coach.py:140)
        acts = np.where(tau_one, sampled, counts.argmax(1))
        nxt, st = tb.rules_step(states, acts.astype(np.int8))
        states[:] = nxt
        ply_no[:] += 1
        for i in np.flatnonzero(st != 0):      # finished game: fresh tree, initial position
            states[i] = 0
            ply_no[i] = 0
            eng.reset(int(i))
        return E * sims

    for _ in range(a.warmup):
        one_ply()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    for _ in range(a.steps):
        done += one_ply()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    dt = reduce_scalar(dt, dist.ReduceOp.MAX if world > 1 else None)
    done = reduce_scalar(float(done), dist.ReduceOp.SUM if world > 1 else None)
    return {"value": done / dt, "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": 1e3 * dt / a.steps, "path": "tc_set_roots_packed -> tc_search -> tc_root_counts -> host move "
            "choice -> tc_rules_step, pinned host buffers"}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
