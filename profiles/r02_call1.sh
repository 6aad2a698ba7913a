mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
./profiles/micro/zipf_l2 > gpurun_out/r2_zipf.log 2>&1
for g in 128 64 32; do
  LGCN_L2_FETCH=$g python profiles/prof_spmm.py amazon plain 6 16 >> gpurun_out/r2_l2fetch.log 2>&1
  LGCN_L2_FETCH=$g python profiles/prof_spmm.py amazon plain 6 32 >> gpurun_out/r2_l2fetch.log 2>&1
done
LGCN_L2_FETCH=64 python profiles/prof_spmm.py amazon plain 6 128 >> gpurun_out/r2_l2fetch.log 2>&1
python profiles/prof_score.py > gpurun_out/r2_score_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_filter -c 1 -o gpurun_out/r2_score_filter python profiles/prof_score.py > gpurun_out/r2_score_ncu.log 2>&1
./profiles/micro/zipf_l2 quick > gpurun_out/r2_zipf_quick.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --csv --log-file gpurun_out/r2_zipf_ncu.csv ./profiles/micro/zipf_l2 quick > gpurun_out/r2_zipf_ncu.log 2>&1
tail -3 gpurun_out/r2_pytest1.log
