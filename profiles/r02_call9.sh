mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -q -x -k "spmm or engine_training or propagate" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
tail -3 gpurun_out/r2_pytest9.log
L=gpurun_out/r2_prefetch.log; : > $L
for f in 0 32; do
  echo "== LGCN_SPMM_FLAGS=$f (32 = no prefetch)" >> $L
  LGCN_SPMM_FLAGS=$f python profiles/prof_adam.py amazon 5 >> $L 2>&1
  LGCN_SPMM_FLAGS=$f python profiles/prof_spmm.py amazon mean 5 >> $L 2>&1
  LGCN_SPMM_FLAGS=$f python profiles/prof_spmm.py amazon add 5 >> $L 2>&1
  LGCN_SPMM_FLAGS=$f python profiles/prof_adam.py amazon 5 16 >> $L 2>&1
  LGCN_SPMM_FLAGS=$f python profiles/prof_spmm.py amazon mean 5 16 >> $L 2>&1
done
grep -E "==|ms=" $L | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench9.json').read()); print(d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
