mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for ev in bf16 bf16_masked; do
  timeout 200 $B --evaluator $ev > gpurun_out/r2y_bench.log 2>&1; echo "$ev rc=$?"
  tail -1 gpurun_out/r2y_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
