#!/usr/bin/env bash
# compute-sanitizer over the kernels whose correctness rests on intra-warp / inter-CTA protocols
# (SURVEY 5 "race detection"): agg_fast_kernel's tag arbitration (store -> __syncwarp -> load), the
# decoupled look-back of filter_kernel and sort_pass_kernel, the peer-exchange flags.  Small sizes:
# the tools slow kernels down 10-100x.  Logs -> gpurun_out/sanitize_*.log, summary -> profiles/.
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh'      (VK_SANITIZE_SEL = pytest -k expression, VK_SANITIZE_TOOLS = tools)
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL=${VK_SANITIZE_SEL:-'test_northstar_filter_aggregate_vs_oracle or test_group_by_hostile_key_distributions and (one_hot_90 or late_groups or sentinel) or test_filter_every_option_vs_numpy_indexing and 2049 or test_sort_every_option_vs_reference or test_sorted_key_column or test_peer_exchange_matches_reference and 2'}
export VK_SANITIZE=1
for tool in ${VK_SANITIZE_TOOLS:-memcheck racecheck initcheck synccheck}; do
  echo "=== $tool ===" | tee -a gpurun_out/sanitize.log
  timeout -k 5 600 $CS --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py tests/test_gpu_dist.py -m gpu -x -q -k "$SEL" \
    > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  echo "rc=$rc ($tool)" | tee -a gpurun_out/sanitize.log
  grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanitize_$tool.log | tail -4 | tee -a gpurun_out/sanitize.log
done
