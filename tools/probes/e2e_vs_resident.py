"""Why a ply of the e2e loop (tc_search from the host, 18.7 ms) is longer than a ply of the device-resident loop (17.95 ms):
per-ply search statistics and CUDA-event time of the search alone, for both drivers on the same staggered positions."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
import tacult_b200 as tb  # noqa: E402
from tacult_b200.weights import random_state_dict  # noqa: E402

E, sims = 4096, 200
eng = tb.Engine(E, sims * 82 + 64)
eng.load_weights(random_state_dict(0, 32, 3, 256))
eng.set_evaluator(tb._lib.EVAL_NET_BF16)
rng = np.random.RandomState(7)
states = np.zeros((E, 8), dtype=np.uint32)
depth = rng.randint(0, 50, size=E)
for d in range(50):
    todo = np.flatnonzero(depth > d)
    mask = tb.rules_legal_mask(states[todo])
    acts = np.array([rng.choice(np.flatnonzero(m)) if m.any() else -1 for m in mask], dtype=np.int8)
    ok = acts >= 0
    nxt, st = tb.rules_step(states[todo][ok], acts[ok])
    keep = st == 0
    states[todo[ok][keep]] = nxt[keep]

def e2e_plies(n):
    global states
    eng.reset(-1)
    out = []
    for ply in range(n):
        eng.set_roots_packed(states)
        eng.reset_stats()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.search(sims, 2.4)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        s = eng.stats()
        counts = eng.root_counts()
        acts = tb.choose_actions(counts, np.zeros(E, bool))
        nxt, st = tb.rules_step(states, acts)
        states = nxt.copy()
        done = np.flatnonzero(st != 0)
        if done.size:
            states[done] = 0
            eng.reset_envs(done)
        out.append((1e3 * dt, s["nn_leaves"] / sims, s["sum_depth"] / max(s["sims"], 1), s["transpositions"] / sims, s["terminal_leaves"] / sims))
    return out

for ply, row in enumerate(e2e_plies(12)):
    print("e2e ply %2d: search %.2f ms, leaves/step %.0f, mean depth %.2f, transpositions/step %.0f, terminal/step %.0f" % ((ply,) + row))
