source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export VINUM_B200_DIST_TRACE=1
TAILN=40 run bench_trace 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --e2e-rows 1000000
