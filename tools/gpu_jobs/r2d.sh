mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_selfplay.py tests/test_gpu_net_bf16.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 6"
for cfg in "TACULT_HALVES=1" "TACULT_HALVES=2" "TACULT_HALVES=3" "TACULT_HALVES=4" "TACULT_HALVES=2 TACULT_PDL=0" "TACULT_HALVES=4 TACULT_PDL=0" "TACULT_HALVES=4 TACULT_FUSED_GMAX=5"; do
  f=gpurun_out/r2d_bench_$(echo $cfg | tr " =" "__").log; env $cfg timeout 200 $B > $f 2>&1; echo "$cfg rc=$?"
  tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],2))"
done
