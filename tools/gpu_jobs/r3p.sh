mkdir -p gpurun_out
t0=$(date +%s.%N); python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err; echo "bench rc=$? wall $(echo "$(date +%s.%N) - $t0" | bc) s"
tail -1 gpurun_out/r3p_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']/1e6,2), d['ms_per_step'], d['e2e']['value'], d['e2e'].get('leaf_evals_per_simulation_step'), d['roofline']['frac'], d['roofline']['rows_per_launch'], d['cpu_baseline']['value'])"
t0=$(date +%s.%N); python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3p_ref.json 2> gpurun_out/r3p_ref.err; echo "ref rc=$? wall $(echo "$(date +%s.%N) - $t0" | bc) s"; tail -1 gpurun_out/r3p_ref.json | cut -c1-200
