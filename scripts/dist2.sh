source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=14 run dist_check 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 1000001
TAILN=4 run bench_n2 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-rows 20000000
grep -o '"ms_per_step": [0-9.]*, "higher\|"step_wall_ms": \[[^]]*\]' gpurun_out/bench_n2.log | head -2
