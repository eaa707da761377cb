"""Where a ply of bench.py's e2e loop spends its host time (C ABI with pinned host buffers, 4096 trees x 200 sims)."""
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
pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
states = pin((E, 8), torch.int32).view(np.uint32)
counts_buf = pin((E, 81), torch.int32).view(np.uint32)
next_buf = pin((E, 8), torch.int32).view(np.uint32)
status_buf = pin((E,), torch.int8)
acts_buf = pin((E,), torch.int8)
ply_no = np.zeros(E, dtype=np.int64)
T = {k: 0.0 for k in ("set_roots", "search_enqueue", "root_counts(sync)", "choose", "rules_step", "resets")}


def one():
    t = time.perf_counter(); eng.set_roots_packed(states); T["set_roots"] += time.perf_counter() - t
    t = time.perf_counter(); eng.search(sims, 2.4); T["search_enqueue"] += time.perf_counter() - t
    t = time.perf_counter(); counts = eng.root_counts(out=counts_buf); T["root_counts(sync)"] += time.perf_counter() - t
    t = time.perf_counter(); acts = tb.choose_actions(counts, (ply_no + 1) < 10, rng.random_sample(E), out=acts_buf); T["choose"] += time.perf_counter() - t
    t = time.perf_counter(); nxt, st = tb.rules_step(states, acts, out=next_buf, status=status_buf); states[:] = nxt; ply_no[:] += 1; T["rules_step"] += time.perf_counter() - t
    t = time.perf_counter()
    done = np.flatnonzero(st != 0)
    if done.size:
        states[done] = 0; ply_no[done] = 0; eng.reset_envs(done)
    T["resets"] += time.perf_counter() - t


for _ in range(30):
    one()
for k in T:
    T[k] = 0.0
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 20
for _ in range(N):
    one()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("ms/ply", round(1e3 * dt / N, 3), {k: round(1e3 * v / N, 3) for k, v in T.items()})
