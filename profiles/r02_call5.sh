mkdir -p gpurun_out
S=gpurun_out/r2_score5.log; : > $S
for v in tc_s3c64 tc_s3c80 tc_s4c64; do echo "== $v" >> $S; LGCN_B200_LIB=profiles/variants/$v.so python profiles/prof_score.py 2>&1 | tail -1 >> $S; done
G=gpurun_out/r2_gowalla_sweep2.log; : > $G
for th in "64 64" "96 96" "48 48" "64 128" "32 64" "48 64" "96 64" "80 80"; do set -- $th
  echo "== threshold $1 seg $2" >> $G
  LGCN_LONG_ROW_THRESHOLD=$1 LGCN_SEG_LEN=$2 python bench.py --workload gowalla --steps 200 --warmup 10 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" >> $G 2>&1
done
cat $S | cut -c1-220; cat $G
bash profiles/r02_sanitizer.sh memcheck
