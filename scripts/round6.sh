source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
# launch order per update: learning chunk (4M rows), then the big chunk -> profile the 2nd agg_fast launch
run ncu_r8 300 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/prof_fast_r8 python scripts/prof_agg.py 8 0
run ncu_r4 300 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/prof_fast_r4 python scripts/prof_agg.py 4 0
