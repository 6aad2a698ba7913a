# 4-GPU lines of the final code (feature sharding, LightGCN and LightGCN_Fusion):  gpurun --gpus 4 -- bash profiles/r02_multigpu4.sh
N=4
mkdir -p gpurun_out
P=29700
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
run bench.py --gpus $N --steps 10 --warmup 3 --parallelism feature > gpurun_out/r2_bench_amazon_n${N}_feature.json 2> gpurun_out/r2_bench_amazon_n${N}_feature.err
run bench.py --gpus $N --steps 10 --warmup 3 --fusion --eval-users 0 > gpurun_out/r2_bench_amazon_n${N}_fusion.json 2> gpurun_out/r2_bench_amazon_n${N}_fusion.err
for f in gpurun_out/r2_bench_amazon_n${N}_*.json; do echo $f; cut -c1-260 $f; done
