mkdir -p gpurun_out
STEADY="python bench.py --steps 1 --warmup 8 --stagger-plies 150 --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_backup_select -s 1800 -c 1 -f -o gpurun_out/prof_k_backup_select_steady_r02_final3 $STEADY > gpurun_out/ncu_k_backup_select_steady_r02_final3.log 2>&1; echo "ncu rc=$?"
