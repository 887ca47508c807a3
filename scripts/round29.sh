source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=60 run sql 900 python -m pytest tests/test_sql_gpu.py -m gpu -q -x --tb=short
