source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
for m in oneshot repartition oneshot repartition; do
export VINUM_B200_DIST_MODE=$m
TAILN=4 run bench_$m 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-rows 1000000
grep -o '"ms_per_step": [0-9.]*, "higher\|"step_wall_ms": \[[^]]*\]' gpurun_out/bench_$m.log | head -2
done
