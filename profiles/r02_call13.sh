mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log; tail -2 gpurun_out/r2_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log; tail -3 gpurun_out/r2_pytest13.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_amazon_final2.json 2> gpurun_out/r2_bench_amazon_final2.err; cut -c1-200 gpurun_out/r2_bench_amazon_final2.json
python bench.py --workload gowalla --steps 200 --warmup 10 > gpurun_out/r2_bench_gowalla_final2.json 2> gpurun_out/r2_bench_gowalla_final2.err; cut -c1-200 gpurun_out/r2_bench_gowalla_final2.json
python bench.py --fusion --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 > gpurun_out/r2_bench_amazon_fusion_1gpu.json 2> gpurun_out/r2_bench_amazon_fusion_1gpu.err; cut -c1-200 gpurun_out/r2_bench_amazon_fusion_1gpu.json
