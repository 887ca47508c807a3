source scripts/gpu_round.sh true
export TAILN=4
run ncu_wide 400 ncu --set full --clock-control none -f --import-source on -k regex:agg_wide -s 1 -c 1 -o gpurun_out/r02_agg_wide_1e6 python scripts/prof_kernels.py groups1e6
