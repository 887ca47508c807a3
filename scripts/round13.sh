source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export VK_CFGS="1:0,0:0"
for v in "" _sync2; do
  export VINUM_B200_LIB=vinum_b200/_C/libvinum_b200$v.so
  TAILN=3 run northstar_parity$v 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or random_types or properties or gtest or fixture"
  TAILN=6 run agg_bench$v 600 python -u scripts/gpu_check.py agg_bench
done
unset VINUM_B200_LIB
run ncu_full 400 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/agg_fast_r02 python scripts/prof_agg.py 1 1000000000
