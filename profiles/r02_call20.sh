mkdir -p gpurun_out
python profiles/prof_fusion.py 1100000 128 > gpurun_out/r2_fusion_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fusion_fwd_tc -c 1 -o gpurun_out/r2_fusion_fwd python profiles/prof_fusion.py 1100000 128 > gpurun_out/r2_fusion_ncu.log 2>&1
cat gpurun_out/r2_fusion_plain.log | tail -8
ncu --set full --clock-control none --import-source on -k regex:fusion_bwd_w_tc -c 1 -o gpurun_out/r2_fusion_bwd python profiles/prof_fusion.py 1100000 128 > gpurun_out/r2_fusion_ncu2.log 2>&1
ls -la gpurun_out/r2_fusion*.ncu-rep
