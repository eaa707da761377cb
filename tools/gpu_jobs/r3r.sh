mkdir -p gpurun_out
echo "== non-streaming (default)"
timeout 900 python -m pytest tests/test_gpu_trunk_col.py tests/test_gpu_net_bf16.py tests/test_gpu_mcts.py -m gpu -x -q 2>&1 | tail -2
python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],3))"
echo "== streaming"
export TACULT_LEAF_STREAM=1
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_selfplay.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],3))"; done
