source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run ncu_c3 400 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/r01_agg_fast_c3 python scripts/prof_c3.py 1000000000
