mkdir -p gpurun_out
TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r3d_timeline.log 2> gpurun_out/r3d_timeline_main.err; echo "rc=$?"
grep "col marks" gpurun_out/r3d_timeline_main.err | tail -1
timeout 900 python -m pytest tests/test_gpu_trunk_col.py tests/test_gpu_net_bf16.py tests/test_gpu_mcts.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],3))"; done
