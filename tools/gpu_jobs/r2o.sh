mkdir -p gpurun_out
for ev in bf16 hash; do
python tools/selfplay_digest.py --evaluator $ev --plies 60 2>&1 | tail -1 | tee -a gpurun_out/r2o_digest.txt
TACULT_B200_LIB=tacult_b200/lib/variants/noprefilter/libtacult_b200.so python tools/selfplay_digest.py --evaluator $ev --plies 60 2>&1 | tail -1 | tee -a gpurun_out/r2o_digest.txt
done
