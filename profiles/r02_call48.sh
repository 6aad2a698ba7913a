# small-graph chunk kernel: warps per CTA / min-blocks variants on the Gowalla step (final plan)
mkdir -p gpurun_out
for v in "" w2 mb6 mb10 w8 ""; do
  if [ -n "$v" ]; then export LGCN_B200_LIB=profiles/variants/$v.so; else unset LGCN_B200_LIB; fi
  python bench.py --workload gowalla --steps 200 --warmup 20 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench48_$v.json 2> gpurun_out/r2_bench48_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench48_$v.json').read().strip().splitlines()[-1])
    print('variant=[$v]', round(d['ms_per_step'],4), {k:round(x['avg_ms'],4) for k,x in d['kernels'].items()})
except Exception as e:
    print('variant=[$v] ERR', e, open('gpurun_out/r2_bench48_$v.err').read()[-400:])
PY
done 2>&1 | tee gpurun_out/r2_small_variants.txt
