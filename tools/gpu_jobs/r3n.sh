mkdir -p gpurun_out
TAG=r02_final3
# steady-state launches: 150 stagger plies @2 sims, 8 warm-up plies @200 sims, then the profiled ply
STEADY="python bench.py --steps 1 --warmup 8 --stagger-plies 150 --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0"
timeout 200 $STEADY > gpurun_out/plain_steady_$TAG.log 2>&1; echo "steady plain rc=$?"; tail -1 gpurun_out/plain_steady_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['rows_per_launch'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6640 --launch-count 560 --csv --log-file gpurun_out/launches_steady_$TAG.csv $STEADY > gpurun_out/ncu_launches_steady_$TAG.log 2>&1; echo "ncu steady launch list rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 6640 --launch-count 560 --csv --log-file gpurun_out/launches_steady_warm_$TAG.csv $STEADY > gpurun_out/ncu_launches_steady_warm_$TAG.log 2>&1; echo "ncu steady warm launch list rc=$?"
for K in k_trunk_col k_backup_select k_fc1_heads; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s 2000 -c 1 -f -o gpurun_out/prof_${K}_steady_$TAG $STEADY > gpurun_out/ncu_${K}_steady_$TAG.log 2>&1; echo "ncu $K rc=$?"
done
