mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02_final3.log 2> gpurun_out/bench_r02_final3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_final3.log 2>&1; echo "reference arm rc=$?"
tail -1 gpurun_out/bench_r02_final3.log | cut -c1-200
