mkdir -p gpurun_out
TACULT_FC1_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r2p.log 2> gpurun_out/r2p.err; echo rc=$?
grep "fc1 marks" gpurun_out/r2p.err | tail -300 | awk 'NR%50==0' | cut -c1-600
