mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -q -x -k "spmm or engine_training" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log; tail -2 gpurun_out/r2_pytest17.log
L=gpurun_out/r2_adam_variants.log; : > $L
for v in main adam_pf adam_b4 adam_pf_b4 main; do
  lib=profiles/variants/$v.so; [ $v = main ] && lib=gcn_recommendation_b200/liblgcn_b200.so
  for d in 128 16; do echo "== $v d=$d" >> $L; LGCN_B200_LIB=$lib python profiles/prof_adam.py amazon 5 $d 2>&1 | tail -1 >> $L; done
done
paste - - < $L | cut -c1-150
