# Round-2 opener.  The end of round 1 measured the opt-in variants once (profiles/r01_variants.md);
# this re-runs the probe sections with the shipped defaults and with each knob off / on, on one box,
# so that the numbers are directly comparable, then the full parity suite.
#   gpurun --timeout 900 -- 'bash scripts/r02_variants.sh'
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export TAILN=3
run defaults 120 python -u scripts/gpu_check.py filter sort arith onegroup topk
# filter: 0 default, 1 = 4096-row tiles, 4 / 8 = batched scatter, 16 / 32 = TMA-staged tiles,
# 64 = 128-thread CTAs, 128 = two row pairs per thread, 192 = both, 256 / 512 = split variant (flags + scan +
# one scatter per column / one scatter launch) -- 64..512 have never run: parity is checked by the section
for c in 256 512 64 128 192 1 4 8 16 32; do
  VINUM_B200_FILTER_CFG=$c run filter_cfg$c 60 python -u scripts/gpu_check.py filter
done
VINUM_B200_CMP_FAST=0 run cmp_off 60 python -u scripts/gpu_check.py filter
VINUM_B200_CMP_FAST=4 run cmp_u4 60 python -u scripts/gpu_check.py filter
VINUM_B200_ARITH_FAST=0 VINUM_B200_ONEGROUP_FAST=0 run arith_onegroup_off 60 python -u scripts/gpu_check.py arith onegroup
VINUM_B200_ARITH_FAST=2 VINUM_B200_ONEGROUP_FAST=2 run arith_onegroup_u2 60 python -u scripts/gpu_check.py arith onegroup
VINUM_B200_ONEGROUP_FUSED=1 run onegroup_fused 60 python -u scripts/gpu_check.py onegroup   # never run: one launch for all functions
VINUM_B200_SORT_FUSE_LAST=1 run sort_fuse_last 60 python -u scripts/gpu_check.py sort   # never run: last pass writes the permutation
VINUM_B200_SORT_FUSE_FIRST=1 run sort_fuse_first 60 python -u scripts/gpu_check.py sort   # never run: first pass computes the codes
VINUM_B200_SORT_FUSE_FIRST=1 VINUM_B200_SORT_FUSE_LAST=1 run sort_fuse_both 60 python -u scripts/gpu_check.py sort
for c in 20 21 22; do   # never run: direct scatter, no shared-memory reorder (16 keys x 3 CTAs, 8 x 5, 8 x 6)
  VINUM_B200_SORT_CFG=$c run sort_cfg$c 60 python -u scripts/gpu_check.py sort
done
VINUM_B200_SORT_PREP=0 run sort_prep_off 60 python -u scripts/gpu_check.py sort
VINUM_B200_SORT_PREP=2 VINUM_B200_TAKE_U=4 run sort_prep_u2_take4 60 python -u scripts/gpu_check.py sort
run pytest_gpu 900 python -m pytest tests -m gpu -x -q
VINUM_B200_ONEGROUP_FUSED=1 VINUM_B200_SORT_FUSE_LAST=1 VINUM_B200_SORT_FUSE_FIRST=1 run pytest_gpu_fused 900 python -m pytest tests -m gpu -x -q
