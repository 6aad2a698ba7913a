# chunk_order on the large-graph ring / live kernels (narrow tables): parity + per-call A/B at the Amazon shape
mkdir -p gpurun_out
python -m pytest tests/test_gpu_config_scale.py -m gpu -q -x -k "neutral" 2>&1 | tail -3
for d in 16 32 64; do for m in plain mean add hop1s hop2; do
  a=$(python profiles/prof_spmm.py amazon $m 6 $d 2>&1 | tail -1 | sed "s/.*ms=//")
  b=$(LGCN_NO_CHUNK_ORDER=1 python profiles/prof_spmm.py amazon $m 6 $d 2>&1 | tail -1 | sed "s/.*ms=//")
  echo "d=$d $m ordered: $a"; echo "d=$d $m natural: $b"
done; done 2>&1 | tee gpurun_out/r2_chunk_order_ab.txt
for d in 16 32; do
  echo "adam d=$d ordered: $(python profiles/prof_adam.py amazon 5 $d 2>&1 | tail -1 | cut -c1-120)"
  echo "adam d=$d natural: $(LGCN_NO_CHUNK_ORDER=1 python profiles/prof_adam.py amazon 5 $d 2>&1 | tail -1 | cut -c1-120)"
done 2>&1 | tee -a gpurun_out/r2_chunk_order_ab.txt
