mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_selfplay.py tests/test_gpu_callers.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for v in "" ""; do
  TACULT_B200_LIB=$v timeout 200 $B > gpurun_out/r2m_bench.log 2>&1; echo "lib=[$v] rc=$?"
  tail -1 gpurun_out/r2m_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
