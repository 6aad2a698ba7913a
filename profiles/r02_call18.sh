mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest18.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest18.log; tail -4 gpurun_out/r2_pytest18.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest18.log | head -20; exit 0; fi
python bench.py --fusion --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench_fusion18.json 2> gpurun_out/r2_bench_fusion18.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_fusion18.json').read()); print('fusion', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, {k:round(v['ms_per_step'],3) for k,v in d['other_kernels'].items()}, d['witness'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench18.json').read()); print('plain', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, d['roofline']['kernel'], d['witness'])"
for d in 16 32; do python profiles/prof_adam.py amazon 5 $d 2>&1 | tail -1; python profiles/prof_spmm.py amazon plain 5 $d 2>&1 | tail -1 | cut -c1-160; done
