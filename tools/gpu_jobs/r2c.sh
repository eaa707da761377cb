mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
timeout 120 ./tools/probes/umma_atmem_probe.bin > gpurun_out/r2c_umma_atmem.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c_umma_atmem.txt
timeout 200 python tools/probes/e2e_prof.py > gpurun_out/r2c_e2e_prof.txt 2>&1; echo "e2e prof rc=$?"; tail -2 gpurun_out/r2c_e2e_prof.txt
timeout 300 python tools/eval_sweep.py --net deep --batches 1024,4096,16384 > gpurun_out/r2c_sweep_deep.jsonl 2>&1; echo "deep rc=$?"; cut -c1-330 gpurun_out/r2c_sweep_deep.jsonl
timeout 300 python tools/eval_sweep.py --net deep --in-planes 32 --batches 1024,4096,16384 > gpurun_out/r2c_sweep_deep_in32.jsonl 2>&1; echo "deep in32 rc=$?"; cut -c1-330 gpurun_out/r2c_sweep_deep_in32.jsonl
