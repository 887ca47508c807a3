source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=8 run sql 900 python -m pytest tests/test_sql_gpu.py -m gpu -q -x --tb=short
TAILN=5 run bench 600 python bench.py --steps 5
