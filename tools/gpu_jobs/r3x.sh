mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3x_pytest.log
t0=$(date +%s); python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3x_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/r3x_smoke.log
python bench.py --skip-cpu-baseline --steps 8 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']/1e6,2), d['e2e']['value'], d['c2_arena'])"
