source scripts/gpu_round.sh true
export TAILN=6
run pytest_all 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
run bench 900 python bench.py
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run ncu_general 400 ncu --set full --clock-control none -f --import-source on -k regex:agg_general -s 1 -c 1 -o gpurun_out/r02_agg_general_1e6 python scripts/prof_kernels.py groups1e6
