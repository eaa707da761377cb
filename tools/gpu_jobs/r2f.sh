mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2f_gpus.txt
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2f_bench_n2.log 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r2f_bench_n2.err
tail -1 gpurun_out/r2f_bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','e2e','c3_full_iteration','c4','trajectory_all_gather','deep_net'): print(k, json.dumps(d.get(k))[:600])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/eval_sweep.py --net deep --batches 2048,8192,32768 > gpurun_out/r2f_sweep_deep_n2.jsonl 2>gpurun_out/r2f_sweep.err; echo "sweep rc=$?"; cut -c1-300 gpurun_out/r2f_sweep_deep_n2.jsonl
