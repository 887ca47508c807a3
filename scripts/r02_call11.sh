source scripts/gpu_round.sh true
export TAILN=4
for pf in 1 2 3; do run filter_pf${pf} 120 python -u scripts/gpu_check.py FILTER_PF=$pf filter; done
for r in 0 1 2; do run sort_rank${r} 120 python -u scripts/gpu_check.py SORT_RANK=$r sort; done
export TAILN=10
run pytest_paths 900 python -m pytest tests/test_gpu_paths.py -m gpu -q --maxfail=20 -p no:cacheprovider -k "filter or sort"
