source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export VK_CFGS="1:0"
for v in "" _k2 _k8; do
  export VINUM_B200_LIB=vinum_b200/_C/libvinum_b200$v.so
  TAILN=3 run northstar_parity$v 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or random_types or properties"
  TAILN=6 run agg_bench$v 600 python -u scripts/gpu_check.py agg_bench
done
unset VINUM_B200_LIB
TAILN=8 run breakdown 300 python scripts/step_breakdown.py
