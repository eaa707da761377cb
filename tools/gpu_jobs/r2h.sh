mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/r2h_bench_n$N.log 2> gpurun_out/r2h_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 400 gpurun_out/r2h_bench_n$N.err
tail -1 gpurun_out/r2h_bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','e2e','c3_full_iteration','c4','trajectory_all_gather','deep_net'): print(k, json.dumps(d.get(k))[:500])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/eval_sweep.py --net deep --batches $((1024*N)),$((4096*N)),$((16384*N)) > gpurun_out/r2h_sweep_deep_n$N.jsonl 2>gpurun_out/r2h_sweep_n$N.err; echo "sweep rc=$?"; cut -c1-260 gpurun_out/r2h_sweep_deep_n$N.jsonl | grep -v NCCL
