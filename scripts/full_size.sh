source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
TAILN=30 run full_size 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or properties_at_scale" --tb=short
