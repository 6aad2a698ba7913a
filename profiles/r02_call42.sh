# chunk_order window size sweep on the narrow large-graph kernels (Amazon shape)
mkdir -p gpurun_out
for W in 32 128 512; do for dm in "16 plain" "16 add" "16 hop2" "16 mean" "32 plain" "32 mean" "32 hop1s" "64 hop1s" "64 mean"; do
  set -- $dm
  a=$(LGCN_CHUNK_ORDER_WINDOW=$W python profiles/prof_spmm.py amazon $2 5 $1 2>&1 | tail -1 | sed "s/.*ms=//" | cut -d, -f3-4)
  echo "W=$W d=$1 $2: $a"
done; done 2>&1 | tee gpurun_out/r2_chunk_order_window.txt
for W in 32 128; do echo "W=$W adam d=16: $(LGCN_CHUNK_ORDER_WINDOW=$W python profiles/prof_adam.py amazon 4 16 2>&1 | tail -1 | cut -c1-100)"; done 2>&1 | tee -a gpurun_out/r2_chunk_order_window.txt
