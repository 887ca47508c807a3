source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=3 run parity 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or random_types or properties"
export VINUM_B200_AGG_PF=4
TAILN=3 run parity_pf4 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or random_types or properties"
unset VINUM_B200_AGG_PF
TAILN=30 run ab 600 python scripts/agg_ab.py "" "AGG_PF=1" "AGG_PF=2" "AGG_PF=4" "AGG_PF=8" "AGG_PF=16" "AGG_PF=4,AGG_WARPS=10" "AGG_PF=4,AGG_WARPS=8" "" 
