mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_config_scale.py "tests/test_gpu_parity.py::test_engine_tail_batch_keeps_the_captured_graph_and_bad_indices_raise" "tests/test_gpu_parity.py::test_topk_ids_and_metrics_vs_reference" -q -x > gpurun_out/r2_pytest4.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest4.log
tail -3 gpurun_out/r2_pytest4.log
if [ $rc -ne 0 ]; then exit 0; fi
S=gpurun_out/r2_score4.log; : > $S
python profiles/prof_score.py 2>&1 | tail -2 >> $S
echo "== tc_s4c64" >> $S; LGCN_B200_LIB=profiles/variants/tc_s4c64.so python profiles/prof_score.py 2>&1 | tail -2 >> $S
L=gpurun_out/r2_relabel.log; : > $L
python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=coldest python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=hot:32000 python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=hot:8000 python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
G=gpurun_out/r2_gowalla_sweep.log; : > $G
for th in "128 64" "64 64" "64 32" "32 32" "32 16" "256 128"; do set -- $th
  echo "== threshold $1 seg $2" >> $G
  LGCN_LONG_ROW_THRESHOLD=$1 LGCN_SEG_LEN=$2 python bench.py --workload gowalla --steps 200 --warmup 10 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" >> $G 2>&1
done
python profiles/prof_score.py > gpurun_out/r2_score_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_filter -c 1 -o gpurun_out/r2_score_filter4 python profiles/prof_score.py > gpurun_out/r2_score_ncu4.log 2>&1
cat $S | cut -c1-200; grep -E "relabel|ms=" $L | cut -c1-200; cat $G
