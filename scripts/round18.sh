source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
TAILN=40 run ab 900 python scripts/agg_ab.py ""
TAILN=5 run bench 600 python bench.py
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --e2e-rows 20000000 --cpu-rows 1000000
run ncu_full 400 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/agg_fast_r03 python scripts/prof_agg.py 1 1000000000
