mkdir -p gpurun_out
L=gpurun_out/r2_adam_narrow.log; : > $L
for d in 16 32 64; do for f in 0 16; do
  echo "== flags=$f d=$d" >> $L; LGCN_SPMM_FLAGS=$f python profiles/prof_adam.py amazon 5 $d 2>&1 | tail -1 >> $L
done; done
paste - - < $L | cut -c1-150
python bench.py --fusion --steps 5 --warmup 3 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fusion', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, d['other_kernels'])"
