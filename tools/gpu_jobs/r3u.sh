mkdir -p gpurun_out
export TACULT_B200_LIB=tacult_b200/lib/variants/unroll1/libtacult_b200.so
for ls in 0 1; do
TACULT_LEAF_STREAM=$ls TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 8 > gpurun_out/r3u_timeline.log 2> gpurun_out/r3u_timeline_$ls.err; echo "stream=$ls rc=$?"
grep "col marks" gpurun_out/r3u_timeline_$ls.err | tail -1 | cut -c1-330
grep "col layers" gpurun_out/r3u_timeline_$ls.err | tail -19 | head -5
TACULT_LEAF_STREAM=$ls timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($ls, round(d['value']/1e6,2), round(d['ms_per_step'],3))"
done
