source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 600 python -m pytest tests -m gpu -x -q
TAILN=40 run agg_parity 300 python -u scripts/gpu_check.py agg_parity
TAILN=30 run agg_bench 600 python -u scripts/gpu_check.py agg_bench
run ncu_d1 300 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/prof_fast_d1 python scripts/prof_agg.py 1
