source scripts/gpu_round.sh true
export VINUM_B200_DEBUG=1
run agg_parity 150 python -u scripts/gpu_check.py agg_parity
VK_HC_ROWS=2000000 run highcard_2m 100 python -u scripts/gpu_check.py agg_highcard
VK_HC_ROWS=20000000 run highcard_20m 150 python -u scripts/gpu_check.py agg_highcard
unset VINUM_B200_DEBUG
VK_TEST_SCALE_ROWS=20000000 TAILN=40 run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run agg_bench 400 python -u scripts/gpu_check.py agg_bench
run smoke 200 python __graft_entry__.py smoke
