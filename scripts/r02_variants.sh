# Round-2 opener: the opt-in kernel variants written (unmeasured, no GPU minutes left) at the end
# of round 1.  Each runs parity first (gpu_check's small cases + the GPU parity tests), then the
# timing at C2 / C4 size; defaults only change once a variant is both green and faster.
#   gpurun --timeout 1500 -- 'bash scripts/r02_variants.sh'
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
# filter_kernel: bit 0 = 4096-row tiles, bits 2-3 = batched scatter (4: single buffer, 8: double);
# filter_tma_kernel: 16 = TMA-staged tiles with 4 column buffers (2 CTAs/SM), 32 = 2 buffers (4 CTAs/SM)
for c in 0 4 8 5 9 16 32; do
  export VINUM_B200_FILTER_CFG=$c
  TAILN=2 run filter_cfg$c 300 python -u scripts/gpu_check.py filter
done
for c in 4 8 16 32; do
  export VINUM_B200_FILTER_CFG=$c
  TAILN=3 run pytest_filter_cfg$c 900 python -m pytest tests -m gpu -x -q -k "filter or where or scale or sql_matches"
done
unset VINUM_B200_FILTER_CFG
# sort_prepare8_kernel<U> and take8_kernel<U>
for u in 0 2 4; do
  export VINUM_B200_SORT_PREP=$u VINUM_B200_TAKE_U=$u
  TAILN=2 run sort_u$u 300 python -u scripts/gpu_check.py sort
done
export VINUM_B200_SORT_PREP=4 VINUM_B200_TAKE_U=4
TAILN=3 run pytest_sort_u4 900 python -m pytest tests -m gpu -x -q -k "sort or order or scale or sql_matches"
unset VINUM_B200_SORT_PREP VINUM_B200_TAKE_U
# compare8_kernel<DOM, U>: plain 8-byte columns (scalar, column-column, BETWEEN)
for u in 0 2 4; do
  export VINUM_B200_CMP_FAST=$u
  TAILN=2 run cmp_fast$u 300 python -u scripts/gpu_check.py filter
done
export VINUM_B200_CMP_FAST=4
TAILN=3 run pytest_cmp_fast4 900 python -m pytest tests -m gpu -x -q -k "compare or between or mask or where or sql_matches"
unset VINUM_B200_CMP_FAST
# arith8_kernel<CC, U>: 8-byte operands in the result's own class
for u in 0 2 4; do
  export VINUM_B200_ARITH_FAST=$u
  TAILN=2 run arith_fast$u 300 python -u scripts/gpu_check.py arith
done
export VINUM_B200_ARITH_FAST=4
TAILN=3 run pytest_arith_fast4 900 python -m pytest tests -m gpu -x -q -k "arith or project or expr or sql_matches"
unset VINUM_B200_ARITH_FAST
# agg_onegroup8_kernel<PK, U>: un-grouped reduction over plain 8-byte columns
for u in 0 2 4; do
  export VINUM_B200_ONEGROUP_FAST=$u
  TAILN=2 run onegroup_fast$u 300 python -u scripts/gpu_check.py onegroup
done
export VINUM_B200_ONEGROUP_FAST=4
TAILN=3 run pytest_onegroup_fast4 900 python -m pytest tests -m gpu -x -q -k "onegroup or one_group or no_group or nogroup or sql_matches or gtest"
unset VINUM_B200_ONEGROUP_FAST
# device top-k (vk_topk_candidates, ops.sort_top; engine opt-in VINUM_B200_TOPK=1) and the rest of
# the code that has never run
VINUM_B200_EXPERIMENTAL=1 TAILN=5 run pytest_experimental 900 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q
TAILN=2 run topk 300 python -u scripts/gpu_check.py topk
