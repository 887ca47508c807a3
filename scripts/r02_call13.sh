source scripts/gpu_round.sh true
export TAILN=6
run sort 120 python -u scripts/gpu_check.py sort
run pytest_all 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
run bench 900 python bench.py
run ncu_filter 300 ncu --set full --clock-control none -f --import-source on -k regex:filter_kernel -s 2 -c 1 -o gpurun_out/r02_filter python scripts/prof_kernels.py filter
run ncu_prep 300 ncu --set full --clock-control none -f -k regex:sort_prepare8 -s 1 -c 1 -o gpurun_out/r02_sort_prepare python scripts/prof_kernels.py sort
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-verify --no-configs
