mkdir -p gpurun_out
TAG=r02_final3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "reference arm rc=$?"
timeout 300 python tools/eval_sweep.py --net trainer --batches 256,1024,4096,16384,65536 > gpurun_out/sweep_trainer_$TAG.jsonl 2>&1; echo "trainer sweep rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --stagger-plies 20 --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0"
timeout 200 $SHORT > gpurun_out/plain_short_$TAG.log 2>&1; echo "short plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm_$TAG.csv $SHORT > gpurun_out/ncu_launches_warm_$TAG.log 2>&1; echo "ncu warm launch list rc=$?"
for K in k_trunk_col k_backup_select k_fc1_heads; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s 300 -c 1 -f -o gpurun_out/prof_${K}_$TAG $SHORT > gpurun_out/ncu_${K}_$TAG.log 2>&1; echo "ncu $K rc=$?"
done
tail -1 gpurun_out/bench_$TAG.log | cut -c1-200
