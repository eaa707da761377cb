mkdir -p gpurun_out
TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r2g_timeline.log 2> gpurun_out/r2g_timeline.err; echo rc=$?
grep "fused marks" gpurun_out/r2g_timeline.err | tail -400 | awk 'NR%40==0' | cut -c1-400
grep "fused timeline" gpurun_out/r2g_timeline.err | tail -9
