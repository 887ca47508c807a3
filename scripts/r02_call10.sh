source scripts/gpu_round.sh true
export TAILN=4
for cs in 0 1; do for pf in 0 1 2; do
  run filter_cs${cs}_pf${pf} 120 python -u scripts/gpu_check.py FILTER_CS=$cs FILTER_PF=$pf filter
done; done
export TAILN=12
run pytest_all 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
run bench 900 python bench.py
