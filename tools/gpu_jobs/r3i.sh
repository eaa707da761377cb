mkdir -p gpurun_out
for s in 0 6 20 40 80; do
python bench.py --skip-cpu-baseline --skip-extras --skip-e2e --settle-plies $s 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print($s, round(d['value']/1e6,2), round(d['ms_per_step'],3), d['roofline']['rows_per_launch'], round(d['roofline']['frac'],3))"
done
