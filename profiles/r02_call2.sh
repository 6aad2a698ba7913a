mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
L=gpurun_out/r2_narrow.log; : > $L
for d in 16 32; do
  echo "== main lib (S_NARROW=8) d=$d default thresholds" >> $L; python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1
  for th in "128 64" "64 32" "256 128"; do set -- $th
    echo "== main lib d=$d threshold $1 seg $2" >> $L; LGCN_LONG_ROW_THRESHOLD=$1 LGCN_SEG_LEN=$2 python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1
  done
  echo "== ring_narrow4 (old) d=$d" >> $L; LGCN_B200_LIB=profiles/variants/ring_narrow4.so python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1
  echo "== ring_narrow16 d=$d" >> $L; LGCN_B200_LIB=profiles/variants/ring_narrow16.so python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1
  echo "== ring_narrow16 d=$d threshold 128 64" >> $L; LGCN_LONG_ROW_THRESHOLD=128 LGCN_SEG_LEN=64 LGCN_B200_LIB=profiles/variants/ring_narrow16.so python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1
done
S=gpurun_out/r2_score_phases.log; : > $S
for p in 1 2 4 8 16 37 148; do LGCN_TC_PHASES=$p python profiles/prof_score.py 2>&1 | tail -2 >> $S; done
for p in 1 8; do echo "== tc_s4c64" >> $S; LGCN_B200_LIB=profiles/variants/tc_s4c64.so LGCN_TC_PHASES=$p python profiles/prof_score.py 2>&1 | tail -2 >> $S; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_amazon.json 2> gpurun_out/r2_bench_amazon.err
python bench.py --workload gowalla --steps 200 --warmup 10 > gpurun_out/r2_bench_gowalla.json 2> gpurun_out/r2_bench_gowalla.err
tail -3 gpurun_out/r2_pytest2.log
