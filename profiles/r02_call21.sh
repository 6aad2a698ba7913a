mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "fusion or guards" > gpurun_out/r2_pytest21.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest21.log; tail -3 gpurun_out/r2_pytest21.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest21.log | head -20; fi
python profiles/prof_fusion.py 4400000 128 2>&1 | tail -6
L=gpurun_out/r2_live_r.log; : > $L
for lib in gcn_recommendation_b200/liblgcn_b200.so profiles/variants/live_r16.so profiles/variants/live_r32.so; do
  for d in 64 128; do for m in hop1s hop2; do
    echo "== $lib" >> $L; LGCN_B200_LIB=$lib python profiles/prof_spmm.py amazon $m 5 $d 2>&1 | tail -1 >> $L
  done; done
done
grep -E "==|ms=" $L | sed 's/N=14700001 nnz=59000000 n_long=2749 n_seg=12834 //' | paste - - | cut -c1-170
