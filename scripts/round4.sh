source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 600 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench 900 python bench.py --steps 5 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
