# The last ~2 GPU minutes of round 1: first contact of the opt-in variants with a GPU, most
# valuable first, every step under its own short timeout (a hang costs 20 s, not the box).
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export TAILN=2
VINUM_B200_FILTER_CFG=4 VINUM_B200_CMP_FAST=4 run ls_filter_batch1 20 python -u scripts/gpu_check.py filter
run ls_filter_default 20 python -u scripts/gpu_check.py filter
VINUM_B200_FILTER_CFG=16 VINUM_B200_CMP_FAST=2 run ls_filter_tma4 20 python -u scripts/gpu_check.py filter
VINUM_B200_FILTER_CFG=8 run ls_filter_batch2 20 python -u scripts/gpu_check.py filter
VINUM_B200_FILTER_CFG=32 run ls_filter_tma2 20 python -u scripts/gpu_check.py filter
VINUM_B200_SORT_PREP=4 VINUM_B200_TAKE_U=4 run ls_sort_u4 25 python -u scripts/gpu_check.py sort
VINUM_B200_ARITH_FAST=4 VINUM_B200_ONEGROUP_FAST=4 run ls_arith_onegroup_fast 25 python -u scripts/gpu_check.py arith onegroup
run ls_arith_onegroup_default 25 python -u scripts/gpu_check.py arith onegroup
run ls_topk 30 python -u scripts/gpu_check.py topk
