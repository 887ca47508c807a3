# 8 GPUs: the multi-rank paths at world 8 + the scaling bench line
source scripts/gpu_round.sh true
export TAILN=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; lscpu | grep -i "^CPU(s)\|NUMA\|Model name" >> gpurun_out/topo_n8.txt
run dist_check_n8 600 $TR scripts/dist_check.py 1000001
run bench_n8 1200 $TR bench.py --gpus 8 --steps 5 --warmup 3
