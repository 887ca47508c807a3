source scripts/gpu_round.sh true
export TAILN=25
run strings 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_sql_gpu.py tests/test_gpu_paths.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "string or generic or more_groups or dropin or sql or fixture"
