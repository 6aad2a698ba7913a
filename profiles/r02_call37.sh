# final single-GPU evidence of round 2 (tests, smoke, three bench lines, launch lists of the final code)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest37.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest37.log; tail -4 gpurun_out/r2_pytest37.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest37.log | head -20; exit 0; fi
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench37_amazon.json 2> gpurun_out/r2_bench37_amazon.err; echo "bench rc=$?"
python bench.py --fusion --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench37_fusion.json 2> gpurun_out/r2_bench37_fusion.err
python bench.py --workload gowalla --steps 200 --warmup 20 > gpurun_out/r2_bench37_gowalla.json 2> gpurun_out/r2_bench37_gowalla.err
for f in amazon fusion gowalla; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench37_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', d.get('e2e',{}).get('value'), {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()}, {k:round(v['ms_per_step'],3) for k,v in d.get('other_kernels',{}).items()}, 'frac', d['roofline']['frac'], d.get('clocks'), (d.get('eval') or {}).get('rating_only_tflops'))
except Exception as e:
    print('$f', 'ERR', e); print(open('gpurun_out/r2_bench37_$f.err').read()[-1500:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches37_fusion.csv python bench.py --fusion --steps 2 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_ncu37a.log 2>&1; echo "ncu fusion rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches37_amazon.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_ncu37b.log 2>&1; echo "ncu amazon rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches37_gowalla.csv python bench.py --workload gowalla --steps 2 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_ncu37c.log 2>&1; echo "ncu gowalla rc=$?"
