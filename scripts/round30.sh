source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
for c in 0 1 2 3 4 5 6; do
  export VINUM_B200_SORT_CFG=$c
  TAILN=2 run sort_cfg$c 300 python -u scripts/gpu_check.py sort
done
