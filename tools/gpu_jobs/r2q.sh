mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rules.py tests/test_gpu_net_bf16.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2q_pytest.log
B="python bench.py --skip-cpu-baseline --skip-extras --steps 8"
for cfg in "TACULT_PDL_TRIGGER=0" "TACULT_PDL_TRIGGER=2" "TACULT_PDL_TRIGGER=0" "TACULT_PDL_TRIGGER=2"; do
  env $cfg timeout 200 $B > gpurun_out/r2q_bench.log 2>&1; echo "$cfg rc=$?"
  tail -1 gpurun_out/r2q_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
