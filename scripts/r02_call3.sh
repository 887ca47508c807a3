source scripts/gpu_round.sh true
export TAILN=14
run pytest_fast 1200 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_gpu_fullsize.py -p no:cacheprovider
run bench 900 python bench.py --no-verify
