source scripts/gpu_round.sh true
export TAILN=12
run skew 900 python -u scripts/skew_probe.py
run bench 900 python bench.py
