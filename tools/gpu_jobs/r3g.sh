mkdir -p gpurun_out
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 8"
for cfg in "TACULT_PDL_TRIGGER=0" "TACULT_PDL_TRIGGER=3" "TACULT_PDL_TRIGGER=4" "TACULT_PDL_TRIGGER=0" "TACULT_PDL_TRIGGER=3" "TACULT_PDL_TRIGGER=4"; do
  f=gpurun_out/r3g_bench_$(echo $cfg | tr " =" "__").log; env $cfg timeout 200 $B > $f 2>&1; echo "$cfg rc=$?"
  tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],3))"
done
