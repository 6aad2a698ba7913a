mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -q -x -k "spmm or engine_training or propagate" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -2 gpurun_out/r2_pytest11.log
L=gpurun_out/r2_mean_adam.log; : > $L
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})" | tee -a $L; }
run main
LGCN_B200_LIB=profiles/variants/meanb2.so run meanb2
LGCN_B200_LIB=profiles/variants/adam_mb6.so run adam_mb6
LGCN_B200_LIB=profiles/variants/adam_mb7.so run adam_mb7
run main
