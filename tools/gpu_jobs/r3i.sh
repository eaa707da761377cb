mkdir -p gpurun_out
python bench.py --skip-cpu-baseline --skip-extras --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), d['e2e'], d['roofline']['rows_per_launch'])"
