mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3f_smoke.log
python bench.py --skip-cpu-baseline --skip-extras --steps 6 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), d['e2e'], d['roofline'], {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
