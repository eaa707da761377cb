mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net_bf16.py -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r3b_pytest.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 6"
for cfg in "TACULT_TRUNK_COL=1" "TACULT_TRUNK_COL=0"; do
  f=gpurun_out/r3b_bench_$(echo $cfg | tr " =" "__").log; env $cfg timeout 200 $B > $f 2>&1; echo "$cfg rc=$?"
  tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r3b_timeline.log 2> gpurun_out/r3b_timeline.err; echo rc=$?
grep "col marks" gpurun_out/r3b_timeline.err | tail -400 | awk 'NR%80==0' | cut -c1-400
