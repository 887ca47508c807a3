# 2 GPUs: the multi-rank paths (peer exchange, sharded Table.sql, sample sort) + bench at N=2
source scripts/gpu_round.sh true
export TAILN=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run dist_check 600 $TR scripts/dist_check.py 2000001
VINUM_B200_DIST_TRACE=1 run bench_n2_trace 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-verify --e2e-rows 20000000
run bench_n2 900 $TR bench.py --gpus 2 --steps 10 --warmup 3
