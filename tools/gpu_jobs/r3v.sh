for v in nopf main nopf main; do
if [ $v = main ]; then unset TACULT_B200_LIB; else export TACULT_B200_LIB=tacult_b200/lib/variants/$v/libtacult_b200.so; fi
timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --skip-extras --steps 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,2), round(d['ms_per_step'],3), {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if 'avg_us' in v})"
done
