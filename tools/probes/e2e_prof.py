import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import tacult_b200 as tb
from tacult_b200.weights import random_state_dict
E, sims = 4096, 200
eng = tb.Engine(E, sims * 82 + 64)
eng.load_weights(random_state_dict(0, 32, 3, 256)); eng.set_evaluator(tb._lib.EVAL_NET_BF16)
rng = np.random.RandomState(7)
states = torch.zeros((E, 8), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
ply_no = np.zeros(E, dtype=np.int64)
T = {k: 0.0 for k in ("set_roots", "search_enqueue", "root_counts(sync)", "choose", "rules_step", "resets")}
def one():
    t = time.perf_counter(); eng.set_roots_packed(states); T["set_roots"] += time.perf_counter() - t
    t = time.perf_counter(); eng.search(sims, 2.4); T["search_enqueue"] += time.perf_counter() - t
    t = time.perf_counter(); counts = eng.root_counts(); T["root_counts(sync)"] += time.perf_counter() - t
    t = time.perf_counter()
    tau_one = (ply_no + 1) < 10
    cum = np.cumsum(counts, axis=1, dtype=np.int64)
    draw = (rng.random_sample(E) * cum[:, -1]).astype(np.int64)
    sampled = (cum > draw[:, None]).argmax(1)
    acts = np.where(tau_one, sampled, counts.argmax(1)); T["choose"] += time.perf_counter() - t
    t = time.perf_counter(); nxt, st = tb.rules_step(states, acts.astype(np.int8)); states[:] = nxt; ply_no[:] += 1; T["rules_step"] += time.perf_counter() - t
    t = time.perf_counter()
    for i in np.flatnonzero(st != 0):
        states[i] = 0; ply_no[i] = 0; eng.reset(int(i))
    T["resets"] += time.perf_counter() - t
for _ in range(30): one()
for k in T: T[k] = 0.0
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 20
for _ in range(N): one()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("ms/ply", 1e3 * dt / N, {k: round(1e3 * v / N, 3) for k, v in T.items()})
