mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3q_smoke.log
