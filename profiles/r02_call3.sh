mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_config_scale.py "tests/test_gpu_parity.py::test_engine_tail_batch_keeps_the_captured_graph_and_bad_indices_raise" "tests/test_gpu_parity.py::test_topk_ids_and_metrics_vs_reference" -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
S=gpurun_out/r2_score3.log; : > $S
python profiles/prof_score.py 2>&1 | tail -2 >> $S
echo "== tc_s4c64" >> $S; LGCN_B200_LIB=profiles/variants/tc_s4c64.so python profiles/prof_score.py 2>&1 | tail -2 >> $S
L=gpurun_out/r2_relabel.log; : > $L
python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=coldest python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=hot:32000 python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
LGCN_RELABEL=hot:8000 python profiles/prof_spmm.py amazon plain 6 128 >> $L 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python profiles/prof_score.py > gpurun_out/r2_score_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_filter -c 1 -o gpurun_out/r2_score_filter3 python profiles/prof_score.py > gpurun_out/r2_score_ncu3.log 2>&1
tail -3 gpurun_out/r2_pytest3.log; cat $S | cut -c1-200; grep -E "relabel|ms=" $L | cut -c1-200
