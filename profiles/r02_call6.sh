mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -4 gpurun_out/r2_pytest6.log
L=gpurun_out/r2_narrow2.log; : > $L
for d in 16 32 128; do python profiles/prof_spmm.py amazon plain 6 $d >> $L 2>&1; done
grep "ms=" $L | cut -c1-200
G=gpurun_out/r2_gowalla_sweep3.log; : > $G
for th in "64 128" "64 256" "48 128" "64 192" "96 128" "32 128"; do set -- $th
  echo "== threshold $1 seg $2" >> $G
  LGCN_LONG_ROW_THRESHOLD=$1 LGCN_SEG_LEN=$2 python bench.py --workload gowalla --steps 200 --warmup 10 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" >> $G 2>&1
done
cat $G
python bench.py --workload gowalla --steps 200 --warmup 10 > gpurun_out/r2_bench_gowalla2.json 2> gpurun_out/r2_bench_gowalla2.err
cut -c1-400 gpurun_out/r2_bench_gowalla2.json
