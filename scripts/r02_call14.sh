source scripts/gpu_round.sh true
export TAILN=6
run pytest_filter 600 timeout 500 python -m pytest tests/test_gpu_paths.py -m gpu -q --maxfail=5 -p no:cacheprovider -k "filter"
export TAILN=4
run filter_base 120 python -u scripts/gpu_check.py filter
run filter_pipe 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 filter
run filter_pipe_pf0 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 FILTER_PF=0 filter
VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_fp6.so run filter_pipe6 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 filter
VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_fp6.so run filter_pipe6_pf0 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 FILTER_PF=0 filter
