source scripts/gpu_round.sh true
TAILN=6
run paths 1500 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "partition or more_groups or every_predicate or general or hostile"
run probe 900 python scripts/groups_probe.py
