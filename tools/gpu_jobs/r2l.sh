mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --stagger-plies 20 --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_backup_select -s 300 -c 1 -f -o gpurun_out/prof_k_backup_select_r2l $SHORT > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
