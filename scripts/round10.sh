# Full evidence pass: gpu tests, smoke, both bench arms, ncu launch list, ncu full on the fused kernel
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt 2>&1
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=5 run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
TAILN=5 run bench 600 python bench.py
TAILN=30 run filter 300 python -u scripts/gpu_check.py filter
TAILN=30 run sort 300 python -u scripts/gpu_check.py sort
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --rows 1000000000 --e2e-rows 20000000 --cpu-rows 1000000
run ncu_full 400 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/agg_fast_r01 python scripts/prof_agg.py 1 1000000000
