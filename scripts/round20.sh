source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=3 run pytest_filter 600 python -m pytest tests -m gpu -x -q -k "filter"
export VINUM_B200_FILTER_PF=0
TAILN=2 run filter_pf0 300 python -u scripts/gpu_check.py filter
export VINUM_B200_FILTER_PF=1
TAILN=2 run filter_pf1 300 python -u scripts/gpu_check.py filter
run ncu_filter 300 ncu --set full --import-source on --clock-control none -k regex:filter_kernel -s 3 -c 1 -f -o gpurun_out/filter_r01 python -u scripts/gpu_check.py filter
