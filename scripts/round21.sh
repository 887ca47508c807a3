source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=3 run pytest_filter 600 python -m pytest tests -m gpu -x -q -k "filter"
TAILN=2 run filter_pf1 300 python -u scripts/gpu_check.py filter
TAILN=2 run sort 300 python -u scripts/gpu_check.py sort
run ncu_sort_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sort.csv python -u scripts/gpu_check.py sort
run ncu_sort 300 ncu --set full --import-source on --clock-control none -k regex:sort_pass -s 12 -c 1 -f -o gpurun_out/sort_pass_r01 python -u scripts/gpu_check.py sort
