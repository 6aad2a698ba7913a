mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
tail -4 gpurun_out/r2_pytest8.log
S=gpurun_out/r2_score8.log; : > $S
python profiles/prof_score.py 2>&1 | tail -1 >> $S
for v in tc_wait100 tc_wait1000 tc_wait20000; do echo "== $v" >> $S; LGCN_B200_LIB=profiles/variants/$v.so python profiles/prof_score.py 2>&1 | tail -1 >> $S; done
cut -c1-200 $S
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_amazon_final.json 2> gpurun_out/r2_bench_amazon_final.err
cut -c1-300 gpurun_out/r2_bench_amazon_final.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_amazon_refarm.json 2> gpurun_out/r2_bench_amazon_refarm.err
cut -c1-400 gpurun_out/r2_bench_amazon_refarm.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_step_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_amazon_step.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_step_ncu.log 2>&1
python profiles/prof_spmm.py amazon plain 4 128 > gpurun_out/r2_ring_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_ring_kernel -s 2 -c 1 -o gpurun_out/r2_spmm_ring_amazon python profiles/prof_spmm.py amazon plain 4 128 > gpurun_out/r2_ring_ncu.log 2>&1
python profiles/prof_spmm.py amazon plain 4 16 > gpurun_out/r2_ring16_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_ring_kernel -s 2 -c 1 -o gpurun_out/r2_spmm_ring_amazon_d16 python profiles/prof_spmm.py amazon plain 4 16 > gpurun_out/r2_ring16_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
