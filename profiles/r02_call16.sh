mkdir -p gpurun_out
L=gpurun_out/r2_adam_narrow2.log; : > $L
for f in 0 16 0 16; do echo "== flags=$f d=128" >> $L; LGCN_SPMM_FLAGS=$f python profiles/prof_adam.py amazon 5 128 2>&1 | tail -1 >> $L; done
paste - - < $L | cut -c1-150
for f in 0 16; do
LGCN_SPMM_FLAGS=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flags=$f', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
