source scripts/gpu_round.sh true
export TAILN=6
run pytest_all 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
run bench 900 python bench.py
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
