mkdir -p gpurun_out
python bench.py --skip-cpu-baseline --steps 8 2>gpurun_out/r3i.err | tail -1 > gpurun_out/r3i_bench.json; python -c "
import json
d=json.load(open('gpurun_out/r3i_bench.json'))
print(round(d['value']/1e6,2), d['e2e']['value'], d['e2e_dropin'], d['c3_full_iteration']['value'])"
tail -3 gpurun_out/r3i.err
