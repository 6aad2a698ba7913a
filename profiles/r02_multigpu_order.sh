# 8-GPU A/B of the per-mode chunk order on narrow tables: bash profiles/r02_multigpu_order.sh 8
N=${1:-8}
mkdir -p gpurun_out
P=29800
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
for v in auto off auto off; do
  LGCN_CHUNK_ORDER_LARGE=$v run bench.py --gpus $N --steps 20 --warmup 3 --parallelism feature --eval-users 0 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_order_$v.json 2> gpurun_out/r2_bench_n${N}_order_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_order_$v.json').read().strip().splitlines()[-1])
print('N=$N order=$v', round(d['ms_per_step'],3), {k:round(x['avg_ms'],3) for k,x in d['kernels'].items()}, d['witness']['param_sum'], d['witness']['top20_ids_crc32'])
PY
done 2>&1 | tee gpurun_out/r2_order_n${N}.txt
