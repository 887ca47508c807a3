source scripts/gpu_round.sh true
export TAILN=4
run ncu_wide 400 ncu --set full --clock-control none -f --import-source on -k regex:agg_wide -s 1 -c 1 -o gpurun_out/r02_agg_wide_1e5 python scripts/prof_kernels.py groups1e5
VINUM_B200_AGG_WIDE=0 run ncu_one 400 ncu --set full --clock-control none -f --import-source on -k regex:agg_general -s 1 -c 1 -o gpurun_out/r02_agg_general_1e5 python scripts/prof_kernels.py groups1e5
