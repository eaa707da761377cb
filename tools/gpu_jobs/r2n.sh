mkdir -p gpurun_out
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for cfg in "TACULT_FC1_HEADS=0 TACULT_FC1_STAGES=3" "TACULT_FC1_HEADS=0 TACULT_FC1_STAGES=4" "TACULT_FC1_HEADS=0 TACULT_FC1_STAGES=2" "TACULT_FC1_HEADS=1"; do
  env $cfg timeout 200 $B > gpurun_out/r2n_bench.log 2>&1; echo "$cfg rc=$?"
  tail -1 gpurun_out/r2n_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
