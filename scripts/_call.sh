source scripts/gpu_round.sh true
TAILN=8
run pytest_all 1800 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
run probe 900 python scripts/groups_probe.py
run bench 900 python bench.py
