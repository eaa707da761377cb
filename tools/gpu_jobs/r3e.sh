mkdir -p gpurun_out
for v in nowaits; do
export TACULT_B200_LIB=tacult_b200/lib/variants/$v/libtacult_b200.so
TACULT_FUSED_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r3e_timeline.log 2> gpurun_out/r3e_timeline_$v.err; echo "$v rc=$?"
grep "col layers" gpurun_out/r3e_timeline_$v.err | tail -10 | head -8 | tail -6
grep "col marks" gpurun_out/r3e_timeline_$v.err | tail -1 | cut -c60-
done
