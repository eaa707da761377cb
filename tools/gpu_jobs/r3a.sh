mkdir -p gpurun_out
for v in 1 2; do
TACULT_FC1_HEADS=$v TACULT_FC1_TIMELINE=1 timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --profile-plies 0 --steps 1 --warmup 1 --stagger-plies 60 > gpurun_out/r3a.log 2> gpurun_out/r3a.err; echo rc=$?
grep "fc1 marks" gpurun_out/r3a.err | tail -1 | cut -c1-600
done
