source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run ncu_sort 300 ncu --set full --import-source on --clock-control none -k regex:sort_pass -s 44 -c 1 -f -o gpurun_out/r01_sort_pass python -u scripts/gpu_check.py sort
run ncu_prep 300 ncu --set full --import-source on --clock-control none -k regex:sort_prepare -s 6 -c 1 -f -o gpurun_out/r01_sort_prepare python -u scripts/gpu_check.py sort
