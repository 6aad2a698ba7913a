# ABI v7 small-graph plan (chunk order, in-launch long-row combine, segments first): parity + Gowalla A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest40.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest40.log; tail -4 gpurun_out/r2_pytest40.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest40.log | head -20; exit 0; fi
for i in 1 2; do
for v in "" 1; do
  if [ -n "$v" ]; then export LGCN_NO_SMALL_PLAN=1; else unset LGCN_NO_SMALL_PLAN; fi
  python bench.py --workload gowalla --steps 200 --warmup 20 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench40_gowalla$v.json 2> gpurun_out/r2_bench40_gowalla$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench40_gowalla$v.json').read().strip().splitlines()[-1])
print('gowalla no_plan=$v', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()}, d['witness']['param_sum'])
PY
done; done 2>&1 | tee gpurun_out/r2_small_plan_ab.txt
unset LGCN_NO_SMALL_PLAN
