mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/r3l_bench_n8.log 2> gpurun_out/r3l_bench_n8.err; echo "bench n8 rc=$?"; tail -c 300 gpurun_out/r3l_bench_n8.err
tail -1 gpurun_out/r3l_bench_n8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','e2e','c3_full_iteration','c4','trajectory_all_gather'): print(k, json.dumps(d.get(k))[:400])"
