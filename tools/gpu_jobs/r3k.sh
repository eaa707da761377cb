mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3k_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r3k_bench_n2.log 2> gpurun_out/r3k_bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/r3k_bench_n2.err
tail -1 gpurun_out/r3k_bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','e2e','c3_full_iteration','c4','trajectory_all_gather'): print(k, json.dumps(d.get(k))[:400])"
