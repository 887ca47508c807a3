source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
TAILN=8 run breakdown 300 python scripts/step_breakdown.py
TAILN=5 run bench 600 python bench.py
TAILN=30 run agg_bench 600 python -u scripts/gpu_check.py agg_bench
