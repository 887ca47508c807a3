source scripts/gpu_round.sh true
export TAILN=8
run paths 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "general or northstar_filter_aggregate"
run probe_small 300 python scripts/groups_probe.py 100000000
run probe 600 python scripts/groups_probe.py
export TAILN=40
VINUM_B200_DEBUG=1 run trace8m 300 python scripts/prof_kernels.py groups8m 1000000000
