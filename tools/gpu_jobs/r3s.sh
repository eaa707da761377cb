mkdir -p gpurun_out
for ls in 1; do
TACULT_LEAF_STREAM=$ls TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 8 > gpurun_out/r3s_timeline.log 2> gpurun_out/r3s_timeline_$ls.err; echo "stream=$ls rc=$?"
grep "col marks" gpurun_out/r3s_timeline_$ls.err | tail -1 | cut -c1-330
grep "col stem" gpurun_out/r3s_timeline_$ls.err | tail -1
grep "col layers" gpurun_out/r3s_timeline_$ls.err | tail -19 | head -9
done
