# small-graph walk with per-tile row / offset preparation (FAST): parity + Gowalla timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest27.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest27.log; tail -4 gpurun_out/r2_pytest27.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest27.log | head -20; exit 0; fi
for m in plain mean add adam; do python profiles/prof_spmm.py gowalla $m 8 2>&1 | tail -1 | cut -c1-260; done
python bench.py --workload gowalla --steps 200 --warmup 20 > gpurun_out/r2_bench27_gowalla.json 2> gpurun_out/r2_bench27_gowalla.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench27_gowalla.json').read().strip().splitlines()[-1])
print('gowalla', round(d['ms_per_step'],4), 'e2e', d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()}, d['witness'])
PY
