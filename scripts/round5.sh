source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export VINUM_B200_DEBUG=1
run smem_prims 200 build/ubench/smem_prims
TAILN=40 run agg_parity 300 python -u scripts/gpu_check.py agg_parity
unset VINUM_B200_DEBUG
TAILN=30 run agg_bench 600 python -u scripts/gpu_check.py agg_bench
