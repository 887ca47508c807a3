source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
TAILN=40 run ab 900 python scripts/agg_ab.py "" ""
TAILN=5 run bench 600 python bench.py
TAILN=5 run bench2 600 python bench.py --steps 30
