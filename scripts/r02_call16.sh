source scripts/gpu_round.sh true
export TAILN=4
run filter_base 120 python -u scripts/gpu_check.py filter
for pf in 1 2 3; do run filter_pipe_pf$pf 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 FILTER_PF=$pf filter; done
run filter_pipe_pf3_cs 120 timeout 100 python -u scripts/gpu_check.py FILTER_PIPE=1 FILTER_PF=3 FILTER_CS=1 filter
