source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_atomic.so
TAILN=3 run parity 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or properties"
TAILN=40 run ab 900 python scripts/agg_ab.py "" "AGG_PF=0,AGG_WARPS=12" "AGG_PF=6,AGG_WARPS=8" "AGG_PF=4,AGG_WARPS=12"
