mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net_bf16.py -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
TACULT_FC1_HEADS=2 timeout 600 python -m pytest tests/test_gpu_net_bf16.py tests/test_gpu_mcts.py -m gpu -q -x -k "not fused_fc1" > gpurun_out/r2z_pytest2.log 2>&1; echo "pytest(v2 default) rc=$?"; tail -2 gpurun_out/r2z_pytest2.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for cfg in "TACULT_FC1_HEADS=1" "TACULT_FC1_HEADS=2" "TACULT_FC1_HEADS=1" "TACULT_FC1_HEADS=2"; do
  env $cfg timeout 200 $B > gpurun_out/r2z_bench.log 2>&1; echo "$cfg rc=$?"
  tail -1 gpurun_out/r2z_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
