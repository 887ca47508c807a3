source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=5 run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
TAILN=5 run bench 600 python bench.py
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --e2e-rows 20000000 --cpu-rows 1000000
