source scripts/gpu_round.sh true
export TAILN=12
run pytest_fast 900 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_gpu_fullsize.py -p no:cacheprovider
run probes 300 python -u scripts/gpu_check.py filter sort onegroup
run probes_nostage 120 python -u scripts/gpu_check.py FILTER_STAGE=0 filter
run pytest_full 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --maxfail=10 -p no:cacheprovider
run bench 900 python bench.py
