# final code after ABI v7: full GPU suite, Gowalla line + launch list, Amazon line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest43.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest43.log; tail -4 gpurun_out/r2_pytest43.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest43.log | head -20; exit 0; fi
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --workload gowalla --steps 200 --warmup 20 > gpurun_out/r2_bench43_gowalla.json 2> gpurun_out/r2_bench43_gowalla.err
python bench.py > gpurun_out/r2_bench43_amazon.json 2> gpurun_out/r2_bench43_amazon.err; echo "bench rc=$?"
for f in amazon gowalla; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench43_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', d.get('e2e',{}).get('value'), 'launches', d['gpu_launches'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()}, 'frac', d['roofline']['frac'], d.get('clocks'), (d.get('eval') or {}).get('rating_only_tflops'))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches43_gowalla.csv python bench.py --workload gowalla --steps 2 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_ncu43c.log 2>&1; echo "ncu gowalla rc=$?"
