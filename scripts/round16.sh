source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
for v in t256k2 t256k4 t256k8; do
export VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_$v.so
TAILN=3 run parity_$v 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "northstar or random_types or properties"
TAILN=40 run ab_$v 900 python scripts/agg_ab.py "AGG_PF=6,AGG_WARPS=8" "AGG_PF=4,AGG_WARPS=8" "AGG_WARPS=8"
done
