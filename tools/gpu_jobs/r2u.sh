mkdir -p gpurun_out
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8"
for v in "" "tacult_b200/lib/variants/trackpos/libtacult_b200.so" "" "tacult_b200/lib/variants/trackpos/libtacult_b200.so"; do
  TACULT_B200_LIB=$v timeout 200 $B > gpurun_out/r2u_bench.log 2>&1; echo "lib=[$v] rc=$?"
  tail -1 gpurun_out/r2u_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
