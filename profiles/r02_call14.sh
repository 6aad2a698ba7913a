mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py tests/test_gpu_tc.py -q -x -k "spmm or engine or amazon_shape or propagate" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
tail -2 gpurun_out/r2_pytest14.log
L=gpurun_out/r2_live_tiles.log; : > $L
for lib in profiles/variants/before_live_tiles.so gcn_recommendation_b200/liblgcn_b200.so; do
  for d in 16 32 64 128; do for m in hop1s hop2; do
    echo "== $lib" >> $L; LGCN_B200_LIB=$lib python profiles/prof_spmm.py amazon $m 5 $d 2>&1 | tail -1 >> $L
  done; done
done
grep -E "==|ms=" $L | sed 's/N=14700001 nnz=59000000 n_long=2749 n_seg=12834 //' | paste - - | cut -c1-200
