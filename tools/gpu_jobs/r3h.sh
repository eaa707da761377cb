mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3h_pytest.log
SHORT="python bench.py --steps 1 --warmup 1 --stagger-plies 20 --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm_r02_final2.csv $SHORT > gpurun_out/ncu_launches_warm.log 2>&1; echo "ncu warm launch list rc=$?"
for n in class16 wide64; do :; done
timeout 300 python tools/eval_sweep.py --net class --batches 4096 2>/dev/null | cut -c1-260
timeout 300 python tools/eval_sweep.py --net wide --batches 4096 2>/dev/null | cut -c1-260
