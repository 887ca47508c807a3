source scripts/gpu_round.sh true
TAILN=25
run newtest 900 python -m pytest tests/test_gpu_paths.py -m gpu -q --maxfail=5 -p no:cacheprovider -k "every_predicate_kind"
TAILN=6
run probe 900 python scripts/groups_probe.py
