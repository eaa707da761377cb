mkdir -p gpurun_out
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 1 --steps 3 --warmup 3"
for h in 2 4; do
TACULT_TRACE=1 TACULT_HALVES=$h timeout 200 $B > gpurun_out/r2e_trace_h$h.log 2> gpurun_out/r2e_trace_h$h.err; echo "h=$h rc=$?"
grep "\[trace\]" gpurun_out/r2e_trace_h$h.err | tail -64 > gpurun_out/r2e_trace_h$h.txt; wc -l gpurun_out/r2e_trace_h$h.txt
done
python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 6 > gpurun_out/r2e_bench.log 2>&1; tail -1 gpurun_out/r2e_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
