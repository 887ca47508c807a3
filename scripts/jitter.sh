source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=6 run ab 300 python scripts/agg_ab.py "" ""
TAILN=8 run breakdown 300 python scripts/step_breakdown.py
