source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=120 run refsuite 600 python -m pytest tests/test_sql_gpu.py -m gpu -q --tb=line -k "reference_suite" -rs
