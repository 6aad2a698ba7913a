mkdir -p gpurun_out
python profiles/prof_fusion.py 1100000 128 > gpurun_out/r2_fusion_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fusion_fwd_tc -c 1 -o gpurun_out/r2_fusion_fwd2 python profiles/prof_fusion.py 1100000 128 > gpurun_out/r2_fusion_ncu3.log 2>&1
tail -5 gpurun_out/r2_fusion_plain2.log
