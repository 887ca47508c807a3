source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
for c in 1 9 8 10 11 12; do
  export VINUM_B200_SORT_CFG=$c
  TAILN=2 run sort_cfg$c 300 python -u scripts/gpu_check.py sort
done
export VINUM_B200_SORT_CFG=9
TAILN=3 run pytest_sort 300 python -m pytest tests -m gpu -x -q -k "sort or sql_matches"
