source scripts/gpu_round.sh true
export TAILN=12
run pytest_agg 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "northstar or hostile or without_predicate or float_keys"
run agg_ab 600 python -u scripts/agg_ab.py
VINUM_B200_AGG_ENTRY=2 run ncu_c3s 400 ncu --set full --clock-control none -f --import-source on -k regex:agg_fast -s 3 -c 1 -o gpurun_out/r02_agg_fast_c3_split python scripts/prof_kernels.py c3
