#!/usr/bin/env bash
# One gpurun call that regenerates everything under profiles/ for a round:
#   gpurun --timeout 2400 -- 'bash scripts/evidence.sh r01'
# (GPU parity tests, smoke, both bench arms, the ncu launch list of the bench command, one
#  ncu --set full capture each of the fused aggregate kernel and of the filter kernel, the
#  launch list of the C4 sort).  Summaries are written by scripts/ncu_summary.py afterwards, here.
R=${1:-r01}
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
TAILN=30 run ab 300 python scripts/agg_ab.py ""
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --e2e-rows 20000000 --cpu-rows 1000000
run ncu_agg 400 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/${R}_agg_fast python scripts/prof_agg.py 1 1000000000
run ncu_filter 300 ncu --set full --import-source on --clock-control none -k regex:filter_kernel -s 3 -c 1 -f -o gpurun_out/${R}_filter python -u scripts/gpu_check.py filter
run ncu_sort_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_sort_launches.csv python -u scripts/gpu_check.py sort
