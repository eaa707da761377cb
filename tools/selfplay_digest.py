#!/usr/bin/env python
"""Plays seeded device-resident self-play and prints a digest of every trajectory record (position, root visit counts,
move, outcome), sorted by (env, game, ply).  Run it with two builds of the library (TACULT_B200_LIB=...) to check that a
kernel variant is bit-identical at scale, e.g. the float32 PUCT pre-filter against the pure float64 selection:

    python tools/selfplay_digest.py --envs 4096 --sims 200 --plies 40
"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import tacult_b200 as tb  # noqa: E402
from tacult_b200.weights import random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--plies", type=int, default=40)
    ap.add_argument("--evaluator", default="bf16", choices=["bf16", "f32", "hash"])
    a = ap.parse_args()
    eng = tb.Engine(a.envs, a.sims * 82 + 64)
    if a.evaluator == "hash":
        eng.set_hash_evaluator(7)
    else:
        eng.load_weights(random_state_dict(0, 32, 3, 256))
        eng.set_evaluator(tb._lib.EVAL_NET_BF16 if a.evaluator == "bf16" else tb._lib.EVAL_NET_F32)
    eng.selfplay_begin(a.sims, 2.4, 10, seed=123, auto_reset=True)
    recs = []
    for _ in range(a.plies):
        eng.selfplay_step(1)
        eng.selfplay_sync()
        recs.append(eng.selfplay_drain())
    recs = np.concatenate(recs)
    order = np.lexsort((recs["ply"], recs["game"], recs["env"]))
    h = hashlib.sha256()
    for name in ("env", "game", "ply", "state", "counts", "action", "z_sign", "finished"):
        h.update(np.ascontiguousarray(recs[name][order]).tobytes())
    st = eng.stats()
    print(f"lib={os.environ.get('TACULT_B200_LIB', 'default')} records={len(recs)} sims={eng.selfplay_stats()['sims']} "
          f"sum_depth={st['sum_depth']} nodes={st['nodes_in_use']} digest={h.hexdigest()}")


if __name__ == "__main__":
    main()
