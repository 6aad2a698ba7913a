mkdir -p gpurun_out
for f in 0 32 0 32; do
  LGCN_SPMM_FLAGS=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flags=$f', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})" | tee -a gpurun_out/r2_prefetch_bench.log
done
