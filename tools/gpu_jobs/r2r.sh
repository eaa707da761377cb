mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2r_smoke.log
