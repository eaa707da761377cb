mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --skip-extras > gpurun_out/r3w_bench_n4.log 2> gpurun_out/r3w_bench_n4.err; echo "bench n4 rc=$?"
tail -1 gpurun_out/r3w_bench_n4.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['ms_per_step'])"
