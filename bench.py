#!/usr/bin/env python
"""Benchmark of the self-play hot path (BASELINE.json metric: MCTS sims/sec, self-play, whole box).

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

A "step" is one lock-step ply of every tree: `sims` simulations per live tree (select -> leaf evaluation
-> expand/backup), move choice from the visit counts, root advance.  Workload at N=1: BASELINE.json
configs[2] — 4096 trees x 200 sims/move, trainer-default ResNet (32 ch x 3 blocks, emb 256), random-init
weights, synthetic self-play.  N>1: every rank runs its own 4096 trees (weak scaling, no collective on the search
path); the `c4` leg then plays BASELINE.json configs[3] — 4096 trees x 800 sims per rank, every game to completion —
and times the iteration's one exchange step (records drained on the device, NCCL all-gather) with it.
Extra keys next to the contract's: `c3_full_iteration` (the reference's iteration shape: no auto-reset, finished
envs idle), `e2e_dropin` (the Python drop-in `VectorizedMCTS.getActionProbs(list[GameState])`), `deep_net`
(configs[4] trunk roofline), `c4` (N>1).  Prints ONE JSON line on rank 0.
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
    ap.add_argument("--evaluator", default="auto", choices=["auto", "f32", "bf16", "bf16_masked"],
                    help="bf16_masked = production mode (masked soft-max in the heads kernel); not the parity-mode headline")
    ap.add_argument("--stagger-plies", type=int, default=150)
    ap.add_argument("--settle-plies", type=int, default=0,
                    help="extra untimed plies at the full simulation count after the stagger (exploration only: the leaf count per "
                         "step moves in waves as games end and restart, and the trunk's two-pass capacity of 3552 leaves is a cliff - "
                         "DESIGN.md section 4 lists value for 0 / 6 / 20 / 40 / 80 settling plies)")
    ap.add_argument("--profile-plies", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-envs", type=int, default=64,
                    help="trees in the host-CPU sample of the reference arm (64 = the reference's own C1 width)")
    ap.add_argument("--c4-sims", type=int, default=800)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="no c3_full_iteration / e2e_dropin / deep_net / c4 legs")
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
# host-CPU arm: the reference's own Python when it is available (/root/reference here, the byte-for-byte staged copy
# under oracle/_ref on the GPU box), else the oracle port
# ------------------------------------------------------------------------------------------------
def live_reference_selfplay(envs, sims, seconds=None, plies=None, warmup_plies=0, seed=0):
    """The UNMODIFIED reference: `Coach.prepExecuteEpisode/executeEpisode` (base/coach.py:87-157) driving its
    `VectorizedMCTS` (base/mcts.py:147-358) over `UtacGame` and `UtacNN` (`_UtacNNet` eager on torch CPU, all host
    threads); the external rules engine `utac_gym` is served by oracle/utac_gym.  Returns None when the reference
    modules are not available."""
    import contextlib
    import io
    import logging
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if not ref_loader.reference_available() and not os.path.isfile(
            os.path.join(ROOT, "oracle", "_ref", "src", "tacult", "base", "mcts.py")):
        return None
    try:
        ref = ref_loader.load_reference()
    except Exception as exc:                      # a dependency of the reference modules is missing on this box
        sys.stderr.write(f"[bench] live reference not loadable ({exc!r}); using the oracle port\n")
        return None
    logging.disable(logging.ERROR)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    np.random.seed(seed)
    args = ref.utils.dotdict(numIters=1, numEnvs=envs, minNumEps=envs, numMCTSSims=sims, cpuct=CPUCT,
                             tempThreshold=TEMP_THRESHOLD, steps_per_epoch=10, epochs=10, batch_size=4096, lr=3e-3,
                             num_warm_restarts=3, onnx_export=True, cuda=False,
                             model_args=dict(dropout=0.2, cuda=False, **NET))
    coach = ref.coach.Coach.__new__(ref.coach.Coach)          # __init__ only adds checkpoint folders and an arena
    coach.game, coach.args, coach.pnet = ref.utac_game.UtacGame(), args, ref.utac_nn.UtacNN(args)
    mcts = ref.mcts.VectorizedMCTS(coach.game, coach.pnet, args)
    coach.prepExecuteEpisode()

    def one_ply():
        live = int((~coach._autoresetEnvs).sum())
        with contextlib.redirect_stdout(io.StringIO()):
            coach.executeEpisode(mcts)
        return live

    for _ in range(warmup_plies):
        one_ply()
    t0 = time.perf_counter()
    done, n, per_ply = 0, 0, []
    while True:
        t1 = time.perf_counter()
        live = one_ply()
        done += live * sims
        per_ply.append((time.perf_counter() - t1) * 1e3)
        n += 1
        if live == 0 or (plies is not None and n >= plies) or (seconds is not None and time.perf_counter() - t0 >= seconds):
            break
    dt = time.perf_counter() - t0
    return done / dt, n, dt, per_ply, os.path.relpath(ref.root, ROOT) if ref.root.startswith(ROOT) else ref.root


def port_selfplay_rate(envs, sims, seconds=None, plies=None, warmup_plies=0, seed=0):
    """Fallback: the reference's CPU path as restated in oracle/ (mcts_oracle.py, game_oracle.py, resnet_oracle.py)."""
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
    over = np.zeros(envs, dtype=bool)

    def one_ply():
        live = int((~over).sum())
        step_no[:] += 1
        canon = [game.getCanonicalForm(boards[i], players[i]) for i in range(envs)]
        probs = mcts.getActionProbs(canon, temps=(step_no < TEMP_THRESHOLD))
        for i in range(envs):
            if over[i]:
                continue
            a = np.random.choice(81, p=probs[i])
            boards[i], players[i] = game.getNextState(boards[i], int(players[i]), a)
            over[i] = game.getGameEnded(boards[i], players[i]) != 0
        return live

    for _ in range(warmup_plies):
        one_ply()
    t0 = time.perf_counter()
    done, n, per_ply = 0, 0, []
    while True:
        t1 = time.perf_counter()
        live = one_ply()
        done += live * sims
        per_ply.append((time.perf_counter() - t1) * 1e3)
        n += 1
        if live == 0 or (plies is not None and n >= plies) or (seconds is not None and time.perf_counter() - t0 >= seconds):
            break
    dt = time.perf_counter() - t0
    return done / dt, n, dt, per_ply, "oracle/"


def host_selfplay_rate(envs, sims, **kw):
    """(kind, sims/s, plies, seconds, per-ply ms, source): the live reference when it can be loaded, else the port."""
    out = live_reference_selfplay(envs, sims, **kw)
    if out is not None:
        return ("reference",) + out
    return ("port",) + port_selfplay_rate(envs, sims, **kw)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    kind, rate, n, dt, per_ply, source = host_selfplay_rate(a.ref_envs, a.sims, plies=a.steps, warmup_plies=a.warmup)
    cores = torch.get_num_threads()
    what = ("the unmodified reference (Coach.executeEpisode -> VectorizedMCTS -> UtacGame -> _UtacNNet, torch CPU eager) "
            f"from {source}, rules engine utac_gym served by oracle/utac_gym") if kind == "reference" else \
        "oracle port of VectorizedMCTS + rules + torch-CPU ResNet"
    sample = (f"{a.ref_envs} of the {a.envs} trees x {a.sims} sims/move from the initial position, {n} plies after "
              f"{a.warmup} warm-up plies in {dt:.1f} s; {what}; the Python search loop is single-threaded, the network "
              f"forward uses {cores} threads")
    line = {
        "impl": "reference", "metric": "mcts_sims_per_sec_selfplay", "value": rate, "unit": "sims/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(per_ply)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": rate, "unit": "sims/s", "cores": cores, "kind": kind, "sample": sample,
                         "sample_trees": a.ref_envs, "host_cpus": os.cpu_count()},
        "e2e": {"value": rate, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a):
    """Identical in both arms (the reference arm times a bounded sample of it, stated in its cpu_baseline.sample)."""
    return {"workload": f"self-play {a.envs} trees x {a.sims} sims/move per GPU, trainer-default ResNet "
                        f"(32ch x 3 blocks, emb 256), random init",
            "trees_per_gpu": a.envs, "sims_per_move": a.sims, "cpuct": CPUCT, "temp_threshold": TEMP_THRESHOLD,
            "tree_arithmetic": "f64",
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
    eng.set_evaluator({"bf16": tb._lib.EVAL_NET_BF16, "bf16_masked": tb._lib.EVAL_NET_BF16_MASKED}.get(evaluator, tb._lib.EVAL_NET_F32))

    # ---- device-resident self-play: the headline `value` --------------------------------------
    seed = 1000 + rank
    eng.selfplay_begin(2, CPUCT, TEMP_THRESHOLD, seed=seed, auto_reset=True, discard_records=True)
    eng.selfplay_step(a.stagger_plies)                       # spread the envs over all game phases
    if a.settle_plies > 0:
        eng.selfplay_configure(sims, CPUCT, TEMP_THRESHOLD, seed=seed, auto_reset=True, discard_records=True)
        eng.selfplay_step(a.settle_plies)                    # trees as they are in the steady state (re-used subtrees)
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

    extras = {}
    if not a.skip_extras:
        # ---- the reference's iteration shape (coach.py:160-193): E games from the initial position to completion, no
        #      auto-reset, finished envs idle (coach.py:114-122,133-134), so the last plies run nearly empty batches ----
        extras["c3_full_iteration"] = full_iteration(eng, torch, stream, sims, seed, barrier, reduce_scalar, dist, world)
        # ---- the Python drop-in under the reference's own calling convention ----
        if rank == 0:
            extras["e2e_dropin"] = run_dropin(a, eng, tb, torch, evaluator)
        barrier()
    eng.close()

    # ---- BASELINE.json configs[4]: the deep 256ch x 20 trunk, where the conv GEMMs are tensor-bound -------------
    if not a.skip_extras:
        extras["deep_net"] = deep_net_leg(tb, torch, stream, local, reduce_scalar, dist, world)
        # ---- BASELINE.json configs[1]: arena of two random-init nets, 1024 games x 2 sims/move (small-batch latency) ----
        if rank == 0:
            extras["c2_arena"] = arena_leg(tb, local)
        barrier()

    # ---- N>1, BASELINE.json configs[3]: 4096 trees x 800 sims per rank, every game to completion, then the iteration's
    #      one exchange step — records drained on the device and all-gathered over NCCL — inside the measured span -----
    gathered = None
    if world > 1 and not a.skip_extras:
        extras["c4"], gathered = c4_leg(a, tb, torch, stream, local, rank, world, evaluator, barrier, reduce_scalar, dist)

    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        kind, rate, n, dt, _, source = host_selfplay_rate(64, 25, seconds=a.cpu_seconds)
        cpu = {"value": rate, "unit": "sims/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"BASELINE config C1 shape: 64 trees x 25 sims/move from the initial position, {n} plies in "
                         f"{dt:.1f} s, " + ("the unmodified reference Python (Coach.executeEpisode -> VectorizedMCTS -> "
                                            f"UtacGame -> _UtacNNet on torch CPU) from {source}, utac_gym served by "
                                            "oracle/utac_gym" if kind == "reference" else
                                            "oracle port of VectorizedMCTS+rules+torch-CPU ResNet")
                         + " (python search loop is single-threaded)",
               "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": "mcts_sims_per_sec_selfplay", "value": value, "unit": "sims/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": evaluator, "data": "synthetic",
            "config": workload_config(a),
            "measured_region": f"steady state: {a.stagger_plies} plies @2 sims with auto-reset spread the envs over all "
                               f"game phases, then {a.settle_plies} settling + {a.warmup} warm-up + {a.steps} timed plies at "
                               f"the full simulation count with auto-reset; the reference's own iteration shape (every "
                               f"transient included) is under c3_full_iteration",
            "leaf_evals_per_sec": leaves * world / (ms * 1e-3),
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roof, "kernels": kernels,
            "kernels_note": "per-kernel times come from a separate pass that brackets every launch with CUDA events; the "
                            "events serialise the programmatic dependent launches, so that pass runs ~10 % slower per ply "
                            "than the headline (shares, not absolutes)",
            "e2e": e2e, "cpu_baseline": cpu, "trajectory_all_gather": gathered,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def full_iteration(eng, torch, stream, sims, seed, barrier, reduce_scalar, dist, world):
    """E games to completion, Σ(live envs x sims) / device time between the first and the last ply."""
    E = eng.num_envs
    eng.selfplay_begin(sims, CPUCT, TEMP_THRESHOLD, seed=seed + 77, auto_reset=False)
    eng.selfplay_sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    k0 = eng.stats()
    ev0.record(stream)
    plies = 0
    while plies < 81:
        eng.selfplay_step(1)
        plies += 1
        if plies >= 17 and eng.selfplay_stats()["games_finished"] >= E:      # no game ends before ply 17
            break
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    eng.selfplay_sync()
    st, k1 = eng.selfplay_stats(), eng.stats()
    ms = reduce_scalar(ev0.elapsed_time(ev1), dist.ReduceOp.MAX if world > 1 else None)
    sims_done = reduce_scalar(float(st["sims"]), dist.ReduceOp.SUM if world > 1 else None)
    records = int(st["examples"])
    eng.selfplay_drain()
    return {"value": sims_done / (ms * 1e-3), "unit": "sims/s", "seconds": ms * 1e-3, "plies": plies,
            "games": int(st["games_finished"]) * world, "env_plies": int(st["plies"]),
            "mean_live_fraction": st["plies"] / float(plies * E), "records": records,
            "leaf_evals_per_sec": (k1["nn_leaves"] - k0["nn_leaves"]) * world / (ms * 1e-3),
            "shape": f"{E} games per GPU from the initial position to completion, {sims} sims/move, no auto-reset "
                     "(base/coach.py:160-193)"}


def all_phase_positions(tb, rng, E):
    """Staggered synthetic positions spread over ALL game phases, the distribution the device-resident loop measures in
    (its envs have been auto-resetting for 150 plies): every env plays one random-move game to its end and starts from a
    uniformly random ply of it.  Returns (packed states u32[E, 8], ply index per env)."""
    cur = np.zeros((E, 8), dtype=np.uint32)
    hist = [cur.copy()]
    alive = np.ones(E, dtype=bool)
    length = np.zeros(E, dtype=np.int64)
    for _ in range(81):
        idx = np.flatnonzero(alive)
        if idx.size == 0:
            break
        mask = tb.rules_legal_mask(cur[idx]).astype(bool)
        acts = (rng.random_sample(mask.shape) * mask).argmax(1).astype(np.int8)      # a uniformly random legal move
        nxt, st = tb.rules_step(cur[idx], acts)
        cur[idx] = nxt
        length[idx] += 1
        alive[idx[st != 0]] = False
        hist.append(cur.copy())
    pick = (rng.random_sample(E) * length).astype(np.int64)          # a ply before the game's last move: never terminal
    return np.stack(hist)[pick, np.arange(E)], pick


def run_dropin(a, eng, tb, torch, evaluator):
    """`tacult_b200.VectorizedMCTS.getActionProbs(list[GameState], temps)` — the call `Coach.executeEpisode` makes
    (coach.py:131) — on 4096 GameState objects x 200 sims: the timed span is the drop-in's own call (host packing of
    the objects, H2D, search, D2H, host tail); moving the boards between calls is the caller's code and is not timed."""
    from tacult_b200.gamestate import GameState
    from tacult_b200.weights import random_state_dict
    E, sims = a.envs, a.sims

    class Args(dict):
        __getattr__ = dict.__getitem__

    class Net:
        class nnet:
            @staticmethod
            def state_dict():
                return random_state_dict(0, **NET)
    args = Args(numEnvs=E, numMCTSSims=sims, cpuct=CPUCT, tc_evaluator=evaluator, tc_max_nodes_per_tree=sims * 82 + 64,
                tc_device=torch.cuda.current_device())
    mcts = tb.VectorizedMCTS(None, Net(), args)
    rng = np.random.RandomState(11)
    states, ply_no = all_phase_positions(tb, rng, E)        # the same distribution as the `e2e` leg
    total_s, pack_s, n_timed = 0.0, 0.0, 0
    np.random.seed(3)
    n_untimed = a.settle_plies + a.warmup
    for it in range(n_untimed + a.steps):
        boards = [GameState.from_packed(s) for s in states]               # the caller's objects
        temps = (ply_no + 1) < TEMP_THRESHOLD
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probs = mcts.getActionProbs(boards, temps)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        tb.pack_gamestates(boards)
        dp = time.perf_counter() - t1
        if it >= n_untimed:
            total_s += dt
            pack_s += dp
            n_timed += 1
        acts = np.array([np.random.choice(81, p=p) for p in probs], dtype=np.int8)       # coach.py:140
        nxt, st = tb.rules_step(states, acts)
        states[:] = nxt
        ply_no += 1
        done = np.flatnonzero(st != 0)
        if done.size:
            states[done] = 0
            ply_no[done] = 0
            mcts.engine.reset_envs(done)
    from tacult_b200.mcts import action_probs_from_counts
    counts = mcts.engine.root_counts().astype(np.int64)
    t2 = time.perf_counter()
    action_probs_from_counts(counts, temps)
    tail_ms = 1e3 * (time.perf_counter() - t2)
    mcts.engine.close()
    return {"value": E * sims * n_timed / total_s, "unit": "sims/s", "ms_per_call": 1e3 * total_s / n_timed,
            "host_pack_ms": 1e3 * pack_s / n_timed, "host_probs_ms": tail_ms,
            "h2d_bytes_per_step": E * 33, "d2h_bytes_per_step": E * 81 * 4,
            "path": "VectorizedMCTS.getActionProbs(list of 4096 tacult_b200.GameState, temps): tc_gs_pack -> "
                    "tc_set_roots_packed -> tc_search -> tc_root_counts -> vectorised counts->probs (global np.random "
                    "draws as the reference's)"}


def deep_net_leg(tb, torch, stream, local, reduce_scalar, dist, world):
    """BASELINE.json configs[4] on this rank: 20 blocks x 256 channels, batch 4096 leaf evaluations resident in HBM;
    algorithmic FLOPs / device time against the measured sustained bf16 peak."""
    from tacult_b200.weights import flops_per_eval, random_state_dict
    C, B, E, batch, iters = 256, 20, 256, 4096, 5
    eng = tb.Engine(batch, 2, device=local)
    eng.set_stream(stream.cuda_stream)
    eng.load_weights(random_state_dict(0, C, B, E))
    res = tb.rules_playouts(1, 0, batch, max_plies=30)
    states = torch.from_numpy(np.array([list(r.final_state) for r in res], dtype=np.uint32).view(np.int32)).cuda()
    pi = torch.empty((batch, 81), dtype=torch.float32, device="cuda")
    v = torch.empty((batch,), dtype=torch.float32, device="cuda")

    def run():
        tb._lib.check(eng.lib.tc_forward_states_dev(eng.handle, tb._lib.EVAL_NET_BF16, batch, states.data_ptr(),
                                                   pi.data_ptr(), v.data_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = reduce_scalar(e0.elapsed_time(e1) / iters, dist.ReduceOp.MAX if world > 1 else None)
    eng.close()
    peaks = measured_peaks()
    tf = flops_per_eval(C, B, E) * batch / (ms * 1e-3) / 1e12
    return {"workload": "256ch x 20 blocks, batch 4096 per GPU, bf16, inputs resident in HBM", "ms": ms,
            "evals_per_sec": batch * world / (ms * 1e-3), "tflops_per_gpu": tf,
            "frac_of_sustained_bf16_peak": tf / peaks["bf16_tflops_sustained"],
            "frac_of_burst_bf16_peak": tf / peaks["bf16_tflops"]}


def arena_leg(tb, local, games=1024, envs=512, sims=2):
    """`VectorizedArena.playGames` semantics (base/arena.py:216-305) on packed boards, host buffers through the C ABI."""
    from tacult_b200.arena import PackedArena
    from tacult_b200.weights import random_state_dict
    engines = []
    for seed in (0, 1):
        e = tb.Engine(envs, 8 * (sims * 82 + 64), device=local)
        e.load_weights(random_state_dict(seed, **NET))
        e.set_evaluator(tb._lib.EVAL_NET_BF16)
        engines.append(e)
    PackedArena(engines[0], engines[1], sims, CPUCT, tie_break="lowest").play_games(2 * envs)       # warm-up pass each way
    arena = PackedArena(engines[0], engines[1], sims, CPUCT, tie_break="lowest")
    t0 = time.perf_counter()
    one, two, draws = arena.play_games(games)
    dt = time.perf_counter() - t0
    ply = np.array(arena.ply_seconds) * 1e3
    for e in engines:
        e.close()
    return {"workload": f"arena {games} games x {sims} sims/move, {envs} boards per pass, two trainer-default nets, bf16",
            "games_per_sec": games / dt, "sims_per_sec": arena.sims / dt, "seconds": dt, "plies": int(len(ply)),
            "ply_ms_p50": float(np.percentile(ply, 50)), "ply_ms_p99": float(np.percentile(ply, 99)),
            "result": [int(one), int(two), int(draws)],
            "path": "host buffers per ply: tc_set_roots_packed -> tc_search -> tc_root_counts -> arg-max -> tc_rules_step"}


def c4_leg(a, tb, torch, stream, local, rank, world, evaluator, barrier, reduce_scalar, dist):
    from tacult_b200.distributed import DeviceReplay, DeviceTrajectoryExchange
    from tacult_b200.weights import random_state_dict
    E, sims = a.envs, a.c4_sims
    eng = tb.Engine(E, sims * 82 + 64, device=local)
    eng.set_stream(stream.cuda_stream)
    eng.load_weights(random_state_dict(0, **NET))
    eng.set_evaluator(tb._lib.EVAL_NET_BF16 if evaluator == "bf16" else tb._lib.EVAL_NET_F32)
    xchg = DeviceTrajectoryExchange(eng, torch.device("cuda", local))
    eng.selfplay_begin(sims, CPUCT, TEMP_THRESHOLD, seed=5000 + rank, auto_reset=False)
    eng.selfplay_sync()
    xchg.warm_up()                                             # receive buffer + NCCL channels up before the clock starts
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    torch.cuda.synchronize()
    ev[0].record(stream)
    plies = 0
    while plies < 81:
        eng.selfplay_step(1)
        plies += 1
        if plies >= 17 and eng.selfplay_stats()["games_finished"] >= E:
            break
    ev[1].record(stream)
    # the ranks finish their last game at different times: align them, so that the exchange is timed on its own and not
    # as "waiting for the slowest rank" (the play phase is already the max over ranks)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    ev[2].record(stream)
    records, counts = xchg.gather()                            # drain on the device + counts + all-gather + compaction
    replay = DeviceReplay(eng, records)
    ev[3].record(stream)
    torch.cuda.synchronize()
    barrier()
    eng.selfplay_sync()
    st = eng.selfplay_stats()
    play_ms = reduce_scalar(ev[0].elapsed_time(ev[1]), dist.ReduceOp.MAX)
    x_ms = reduce_scalar(ev[2].elapsed_time(ev[3]), dist.ReduceOp.MAX)
    all_ms = play_ms + x_ms
    sims_done = reduce_scalar(float(st["sims"]), dist.ReduceOp.SUM)
    n_total = int(sum(counts))
    # a learner-side use of the gathered records, still on the device: one batch of examples (8 symmetries on the fly)
    idx = torch.randint(0, max(replay.n_examples, 1), (4096,), device="cuda", dtype=torch.int32)
    obs, pi, z = replay.examples(idx)
    ok = bool(torch.isfinite(obs).all() and torch.isfinite(pi).all() and (pi.sum(1) - 1).abs().max() < 1e-3)
    gathered = {"records_local": int(counts[rank]), "records_all_ranks": n_total, "ms": x_ms, "bytes_per_record": 216,
                "effective_gbps": n_total * 216 / (x_ms * 1e-3) / 1e9 if x_ms > 0 else None,
                "path": "tc_selfplay_drain_dev -> persistent CUDA send buffer -> all_gather_into_tensor (NCCL) -> "
                        "device record tensor -> tc_examples_from_records_dev", "examples_ok": ok}
    eng.close()
    return ({"workload": f"BASELINE configs[3]: {world} x {E} trees x {sims} sims/move, every game to completion, records "
                         "all-gathered over NCCL",
             "value_with_exchange": sims_done / (all_ms * 1e-3), "value_without_exchange": sims_done / (play_ms * 1e-3),
             "unit": "sims/s", "seconds_play": play_ms * 1e-3, "seconds_exchange": x_ms * 1e-3, "plies": plies,
             "games": int(st["games_finished"]) * world, "records": n_total,
             "exchange_share": x_ms / all_ms if all_ms else None}, gathered)


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
        # the tensor pipe of a 32-channel net is fed by shared-memory operand reads at 128 B/clk, not by its math rate
        roof["note"] = ("column-tiled trunk k_trunk_col: the three horizontal taps stacked in N (6 UMMAs of N = 96 per tile and layer "
                        "instead of 18 + 3 of N = 32); bound by shared-memory bandwidth - 42 KB of operand fetch + 8-16 KB of epilogue "
                        "rows per tile and layer through 128 B/clk = 390-450 cycles, measured 428 (DESIGN.md section 4)")
    return roof, kernels


def run_e2e(a, eng, tb, torch, rank, barrier, reduce_scalar, dist, world):
    """Same metric through the C ABI with HOST buffers: every ply copies the root positions host->device,
    searches, reads the visit counts device->host, picks moves on the host and advances the positions with
    tc_rules_step (host buffers again).  Inputs live in pinned host memory."""
    E, sims = a.envs, a.sims
    rng = np.random.RandomState(7 + rank)
    states = torch.zeros((E, 8), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    # results land in pinned host buffers too (visit counts, successor positions, their status)
    counts_buf = torch.zeros((E, 81), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    next_buf = torch.zeros((E, 8), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    status_buf = torch.zeros((E,), dtype=torch.int8).pin_memory().numpy()
    acts_buf = torch.zeros((E,), dtype=torch.int8).pin_memory().numpy()
    # Positions over all game phases (all_phase_positions).  Until round 2's last bench this was "random moves to a random
    # depth below 50", which has no late-game positions: 3.6-3.9 k leaves per simulation instead of 3.0-3.4 k, i.e. beyond
    # the 3552 leaves the trunk evaluates in two passes per SM - a different workload from the one `value` is quoted on.
    states[:], ply_no = all_phase_positions(tb, rng, E)
    eng.reset(-1)
    eng.reset_stats()
    h2d = E * 32 + E * 32 + E          # roots + positions and actions handed to tc_rules_step
    d2h = E * 81 * 4 + E * 32 + E      # visit counts + successor positions and their status

    def one_ply():
        eng.set_roots_packed(states)
        eng.search(sims, CPUCT)
        counts = eng.root_counts(out=counts_buf)
        # host move choice, one C call (tc_choose_actions): tau = 1 -> a ~ counts / sum (coach.py:140), tau = 0 -> arg-max
        tau_one = (ply_no + 1) < TEMP_THRESHOLD
        acts = tb.choose_actions(counts, tau_one, rng.random_sample(E), out=acts_buf)
        nxt, st = tb.rules_step(states, acts, out=next_buf, status=status_buf)
        states[:] = nxt
        ply_no[:] += 1
        done = np.flatnonzero(st != 0)         # finished games: fresh trees (VectorizedMCTS.reset), initial position
        if done.size:
            states[done] = 0
            ply_no[done] = 0
            eng.reset_envs(done)
        return E * sims

    for _ in range(a.settle_plies + a.warmup):            # settle: the same untimed plies as the device-resident loop's
        one_ply()
    k0 = eng.stats()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    for _ in range(a.steps):
        done += one_ply()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    k1 = eng.stats()
    leaves_per_step = (k1["nn_leaves"] - k0["nn_leaves"]) / max(1, a.steps * sims)
    dt = reduce_scalar(dt, dist.ReduceOp.MAX if world > 1 else None)
    done = reduce_scalar(float(done), dist.ReduceOp.SUM if world > 1 else None)
    return {"value": done / dt, "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": 1e3 * dt / a.steps, "path": "tc_set_roots_packed -> tc_search -> tc_root_counts -> host move "
            "choice (tc_choose_actions) -> tc_rules_step, pinned host buffers",
            "positions": "every env starts from a uniformly random ply of a random-move game (all game phases, as in the "
                         "device-resident loop); finished games restart from the initial position",
            "leaf_evals_per_simulation_step": leaves_per_step}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
