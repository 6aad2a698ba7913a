# compute-sanitizer over the smallest cases that reach every kernel family (SURVEY.md section 5,
# VERDICT r01 item 7).  ONE tool per gpurun call:  bash profiles/r02_sanitizer.sh memcheck|racecheck
tool=${1:-memcheck}
mkdir -p gpurun_out
sel='tests/test_gpu_parity.py::test_engine_training_matches_reference[tiny_lightgcn_d64_k3-False]
tests/test_gpu_parity.py::test_engine_training_matches_reference[tiny_fusion_d64_k3-False]
tests/test_gpu_parity.py::test_engine_brand_bpr_term_matches_reference[False]
tests/test_gpu_parity.py::test_spmm_large_graph_kernels_bit_exact[ring-16]
tests/test_gpu_parity.py::test_spmm_large_graph_kernels_bit_exact[ring_hot-128]
tests/test_gpu_parity.py::test_spmm_large_graph_kernels_bit_exact[chunk-64]
tests/test_gpu_parity.py::test_topk_ids_and_metrics_vs_reference[tiny_lightgcn_d64_k3]
tests/test_gpu_parity.py::test_device_sampler_epoch_is_a_permutation_with_valid_negatives
tests/test_gpu_tc.py::test_tc_topk_matches_exact_kernel[20000-128]
tests/test_gpu_tc.py::test_tc_batched_sweep_and_item_splits_match_exact_kernel'
# plain run first (exit code visible), then the same selection under the tool
python -m pytest -q -x $sel > gpurun_out/r2_san_plain.log 2>&1 || { tail -5 gpurun_out/r2_san_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/r2_san_$tool.log \
  python -m pytest -q -x $sel > gpurun_out/r2_san_${tool}_pytest.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2_san_${tool}_pytest.log
tail -3 gpurun_out/r2_san_${tool}_pytest.log; tail -5 gpurun_out/r2_san_$tool.log
