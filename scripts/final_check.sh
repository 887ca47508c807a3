# Last gate of a round: the driver's own sequence (GPU tests, smoke, both bench arms).
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=5 run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
TAILN=5 run bench 600 python bench.py
