mkdir -p gpurun_out
TACULT_FUSED_BALANCE=1 timeout 600 python -m pytest tests/test_gpu_net_bf16.py -m gpu -q -x > gpurun_out/r2w_pytest.log 2>&1; echo "pytest(balance) rc=$?"; tail -1 gpurun_out/r2w_pytest.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for cfg in "TACULT_FUSED_BALANCE=0" "TACULT_FUSED_BALANCE=1" "TACULT_FUSED_BALANCE=0" "TACULT_FUSED_BALANCE=1"; do
  env $cfg timeout 200 $B > gpurun_out/r2w_bench.log 2>&1; echo "$cfg rc=$?"
  tail -1 gpurun_out/r2w_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
