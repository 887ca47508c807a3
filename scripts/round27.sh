source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=60 run generic 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "generic"
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
