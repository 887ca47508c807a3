source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
nvidia-smi -L > gpurun_out/gpus8.txt
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TAILN=12 run dist_check8 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
TAILN=4 run bench_n8 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3
TAILN=4 run bench_n4 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3
