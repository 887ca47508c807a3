source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
