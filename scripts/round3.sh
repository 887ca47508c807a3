source scripts/gpu_round.sh true
run ncu_s0 300 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/prof_agg_s0 python scripts/prof_agg.py 0 12
run ncu_s1 300 ncu --set full --import-source on --clock-control none -k regex:agg_fast -s 1 -c 1 -f -o gpurun_out/prof_agg_s1 python scripts/prof_agg.py 1 11
run ncu_filter 300 ncu --set full --import-source on --clock-control none -k regex:filter_kernel -s 2 -c 1 -f -o gpurun_out/prof_filter python scripts/gpu_check.py filter
run ncu_sort 300 ncu --set full --import-source on --clock-control none -k regex:sort_pass -s 12 -c 1 -f -o gpurun_out/prof_sort python scripts/gpu_check.py sort
ls -la gpurun_out/
