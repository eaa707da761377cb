#!/usr/bin/env python
"""BASELINE.json configs[1] / SURVEY.md §8d C2: arena of 1024 games x 2 sims per move, 512 boards per pass, two
random-init trainer-default nets (seeds 0 and 1) in bf16, tau = 0.  Reports games/s, sims/s and per-ply latency.

    python tools/arena_bench.py [--games 1024] [--envs 512] [--sims 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import tacult_b200 as tb  # noqa: E402
from tacult_b200.arena import PackedArena  # noqa: E402
from tacult_b200.weights import random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1024)
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--sims", type=int, default=2)
    ap.add_argument("--evaluator", default="bf16", choices=["bf16", "f32"])
    a = ap.parse_args()
    ev = tb._lib.EVAL_NET_BF16 if a.evaluator == "bf16" else tb._lib.EVAL_NET_F32
    engines = []
    for seed in (0, 1):
        e = tb.Engine(a.envs, 8 * (a.sims * 82 + 64))
        e.load_weights(random_state_dict(seed, 32, 3, 256))
        e.set_evaluator(ev)
        engines.append(e)
    warm = PackedArena(engines[0], engines[1], a.sims, 2.4, tie_break="lowest")
    warm.play_games(2 * a.envs)                       # one pass each way: warm-up (trees persist, as in the reference)
    arena = PackedArena(engines[0], engines[1], a.sims, 2.4, tie_break="lowest")
    t0 = time.perf_counter()
    one, two, draws = arena.play_games(a.games)
    dt = time.perf_counter() - t0
    ply = np.array(arena.ply_seconds) * 1e3
    print(json.dumps({"workload": f"arena {a.games} games x {a.sims} sims/move, {a.envs} boards per pass, two trainer-default nets",
                      "evaluator": a.evaluator, "seconds": dt, "games_per_sec": a.games / dt, "sims_per_sec": arena.sims / dt,
                      "plies": len(ply), "ply_ms_p50": float(np.percentile(ply, 50)), "ply_ms_p99": float(np.percentile(ply, 99)),
                      "result": [one, two, draws],
                      "path": "host buffers: tc_set_roots_packed -> tc_search -> tc_root_counts -> argmax -> tc_rules_step"}))


if __name__ == "__main__":
    main()
