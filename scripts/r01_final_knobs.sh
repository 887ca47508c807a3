# Full GPU parity suite under the knobs that won the first-contact probe, plus the gated tests of
# the never-run code, side by side (two processes) to fit the round's last GPU seconds.
mkdir -p gpurun_out
export VINUM_B200_ARITH_FAST=4 VINUM_B200_ONEGROUP_FAST=4 VINUM_B200_CMP_FAST=2 VINUM_B200_SORT_PREP=4 VINUM_B200_TOPK=1
(timeout 42 python -m pytest tests -m gpu -x -q > gpurun_out/final_knobs_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/final_knobs_suite.log) &
(VINUM_B200_EXPERIMENTAL=1 timeout 42 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q > gpurun_out/final_experimental.log 2>&1; echo "experimental rc=$?" >> gpurun_out/final_experimental.log) &
wait
tail -4 gpurun_out/final_knobs_suite.log
tail -12 gpurun_out/final_experimental.log | cut -c1-400
