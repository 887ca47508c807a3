source scripts/gpu_round.sh true
export TAILN=4
VINUM_B200_FILTER_PIPE=1 run ncu_pipe 300 ncu --set full --clock-control none -f --import-source on -k regex:filter_pipe -s 2 -c 1 -o gpurun_out/r02_filter_pipe python scripts/prof_kernels.py filter
