# ALT as a template parameter: parity (full GPU suite) + A/B against the variant with the override compiled out
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest36.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest36.log; tail -4 gpurun_out/r2_pytest36.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest36.log | head -20; exit 0; fi
for m in plain mean add; do echo "main $m: $(python profiles/prof_spmm.py amazon $m 6 2>&1 | tail -1 | sed "s/.*ms=//")"; echo "noalt $m: $(LGCN_B200_LIB=profiles/variants/spmm_noalt.so python profiles/prof_spmm.py amazon $m 6 2>&1 | tail -1 | sed "s/.*ms=//")"; done 2>&1 | tee gpurun_out/r2_alt_ab2.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench36.json 2> gpurun_out/r2_bench36.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench36.json').read()); print('amazon', round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, d['roofline']['frac'], d['clocks'])"
