# multi-GPU evidence (run with gpurun --gpus N):  bash profiles/r02_multigpu.sh N
N=${1:-2}
mkdir -p gpurun_out
P=29500
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
# 1. real-NCCL parity of the sharded engines (tests/dist_gpu_worker.py)
run tests/dist_gpu_worker.py > gpurun_out/r2_dist_parity_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_parity_n$N.log
# 2. bench lines with the cross-N witness: feature sharding (default) and the north-star row sharding
run bench.py --gpus $N --steps 10 --warmup 3 --parallelism feature > gpurun_out/r2_bench_amazon_n${N}_feature.json 2> gpurun_out/r2_bench_amazon_n${N}_feature.err
run bench.py --gpus $N --steps 10 --warmup 3 --parallelism row --eval-users 131072 > gpurun_out/r2_bench_amazon_n${N}_row.json 2> gpurun_out/r2_bench_amazon_n${N}_row.err
if [ "$2" = "fusion" ]; then
run bench.py --gpus $N --steps 10 --warmup 3 --fusion --eval-users 0 > gpurun_out/r2_bench_amazon_n${N}_fusion.json 2> gpurun_out/r2_bench_amazon_n${N}_fusion.err
fi
tail -2 gpurun_out/r2_dist_parity_n$N.log; for f in gpurun_out/r2_bench_amazon_n${N}_*.json; do echo $f; cut -c1-300 $f; done
