mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import tacult_b200 as tb
from tacult_b200.weights import random_state_dict
eng = tb.Engine(256, 40 * 82 + 64)
eng.load_weights(random_state_dict(0, 32, 3, 256))
eng.set_evaluator(tb._lib.EVAL_NET_BF16)
eng.selfplay_begin(24, 2.4, 10, seed=3, auto_reset=True)
eng.selfplay_step(12)
eng.selfplay_sync()
print("records", len(eng.selfplay_drain()), eng.selfplay_stats())
eng2 = tb.Engine(64, 30 * 82 + 64, hash_salt=5)
eng2.set_roots_packed(np.zeros((64, 8), dtype=np.uint32))
eng2.search(30, 2.4)
print("hash counts", int(eng2.root_counts().sum()))
print(tb.rules_playouts(1, 0, 512)[0].checksum)
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2v_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2v_memcheck.log
