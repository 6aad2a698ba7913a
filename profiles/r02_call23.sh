mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "fusion" > gpurun_out/r2_pytest23.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2_pytest23.log; tail -3 gpurun_out/r2_pytest23.log
if [ $rc -ne 0 ]; then grep -n "Error\|^E " gpurun_out/r2_pytest23.log | head -20; fi
for lib in profiles/variants/fusion_old_loader.so gcn_recommendation_b200/liblgcn_b200.so profiles/variants/fusion_old_loader.so gcn_recommendation_b200/liblgcn_b200.so; do
  echo "== $lib"; LGCN_B200_LIB=$lib python profiles/prof_fusion.py 4400000 128 2>&1 | grep -E " tc |rel err"
done
LGCN_B200_LIB=gcn_recommendation_b200/liblgcn_b200.so python profiles/prof_fusion.py 4400000 64 2>&1 | grep -E " tc |rel err"
