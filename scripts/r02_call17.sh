source scripts/gpu_round.sh true
export TAILN=6
run pytest_agg 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py tests/test_gpu_dist.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "northstar or hostile or without_predicate or float_keys or peer"
run agg_ab 600 python -u scripts/agg_ab.py
run skew 900 python -u scripts/skew_probe.py
