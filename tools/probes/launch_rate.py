import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import tacult_b200 as tb
for E in (64, 4096):
    eng = tb.Engine(E, 400 * 82 + 64, hash_salt=3)
    eng.set_roots_packed(np.zeros((E, 8), dtype=np.uint32))
    eng.search(50, 2.4); eng.root_counts()
    t0 = time.perf_counter(); eng.search(300, 2.4); t1 = time.perf_counter(); eng.root_counts(); t2 = time.perf_counter()
    print(f"E={E}: enqueue of 600 launches {1e3*(t1-t0):.2f} ms ({1e6*(t1-t0)/600:.1f} us per launch), until done {1e3*(t2-t0):.2f} ms")
