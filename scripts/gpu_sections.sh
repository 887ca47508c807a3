#!/usr/bin/env bash
# Run each probe section in its own process with its own timeout; logs stream into
# gpurun_out/ so a hang in one section cannot hide the others.
mkdir -p gpurun_out
T=${SECTION_TIMEOUT:-120}
for s in "$@"; do
  echo "=== $s ===" | tee -a gpurun_out/sections.log
  timeout -k 5 "$T" python -u scripts/gpu_check.py "$s" > "gpurun_out/check_$s.log" 2>&1
  rc=$?
  echo "rc=$rc" | tee -a gpurun_out/sections.log
  tail -12 "gpurun_out/check_$s.log" | cut -c1-1500
done
