#!/usr/bin/env bash
# One gpurun call, many answers.  Every step has its own timeout and log under gpurun_out/.
mkdir -p gpurun_out
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name ===" | tee -a gpurun_out/round.log
  timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "rc=$? ($name)" | tee -a gpurun_out/round.log
  tail -${TAILN:-15} "gpurun_out/$name.log" | cut -c1-1600
}
"$@"
