mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_callers.py tests/test_gpu_selfplay.py tests/test_gpu_mcts.py -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log
timeout 400 python bench.py --skip-cpu-baseline --steps 8 > gpurun_out/r2t_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r2t_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'dropin', round(d['e2e_dropin']['value']/1e6,2), d['e2e_dropin']['ms_per_call'], 'full', round(d['c3_full_iteration']['value']/1e6,2))"
