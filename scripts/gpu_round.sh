#!/usr/bin/env bash
# One gpurun call, many answers.  Every step has its own timeout and log under gpurun_out/.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh validate'            tests + bench + smoke (what the driver runs)
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh profile'             one ncu --set full capture per kernel + the launch list
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh probes'              the A/B tables of profiles/r02_tuning.md, r02_skew.md
#   gpurun --gpus 2 --timeout 1800 -- 'bash scripts/gpu_round.sh dist 2'     multi-GPU parity + bench at N ranks
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh sanitize'            compute-sanitizer passes (scripts/sanitize.sh)
# Summaries for profiles/: python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_ncu_full.md
mkdir -p gpurun_out
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name ===" | tee -a gpurun_out/round.log
  timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "rc=$? ($name)" | tee -a gpurun_out/round.log
  tail -${TAILN:-15} "gpurun_out/$name.log" | cut -c1-1600
}

validate() {
  TAILN=6
  run pytest_all 1800 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider
  run bench 900 python bench.py
  run bench_reference 900 python bench.py --impl reference
  run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
}

# One ncu --set full capture, summarised on the box: only the .md comes back (a dozen reports with source
# exceed what gpurun merges back); KEEP_REP=1 keeps the report as well.
cap() { # name kernel-regex launches-to-skip workload
  local name=$1 rex=$2 skip=$3 what=$4
  run "ncu_$name" 400 ncu --set full --clock-control none -f --import-source on -k "regex:$rex" -s "$skip" -c 1 \
      -o "gpurun_out/r02_$name" python scripts/prof_kernels.py "$what"
  python scripts/ncu_summary.py "gpurun_out/r02_$name.ncu-rep" > "gpurun_out/r02_${name}_ncu_full.md" 2> "gpurun_out/ncu_summary_$name.err"
  [ -n "$KEEP_REP" ] || rm -f "gpurun_out/r02_$name.ncu-rep"
}

profile() {
  TAILN=2
  cap agg_fast agg_fast 3 northstar
  cap agg_fast_dict agg_fast 3 hash
  cap agg_fast_c3 agg_fast 3 c3
  # 2e8 rows = launches of 7.2 M, 65 M and 127 M rows per aggregate (chunks grow while the table is sized): the last one
  cap agg_wide_1e6 agg_wide 5 groups1e6
  VINUM_B200_AGG_WIDE=0 cap agg_general_1e6 agg_general 5 groups1e6
  # 8.4e6 groups: the partitioned plan's two kernels, third round of the first aggregate (all groups exist)
  cap agg_part_scatter_8m agg_part_scatter 2 groups8m
  cap agg_part_update_8m agg_part_update 2 groups8m
  cap filter filter_kernel 2 filter
  cap sort_pass sort_pass 10 sort
  cap sort_prepare sort_prepare8 1 sort
  cap arith8 arith8 2 arith
  cap compare8 compare8 2 compare
  cap onegroup8 agg_onegroup8 4 onegroup
  cap topk_hist topk_hist 2 topk
  run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
      python bench.py --steps 2 --warmup 1 --no-verify --no-configs
}

probes() {
  TAILN=12
  run agg_ab 600 python -u scripts/agg_ab.py
  run groups 600 python -u scripts/groups_probe.py
  run skew 900 python -u scripts/skew_probe.py
  run hot 300 python -u scripts/hot_probe.py
  run ingest 300 python -u scripts/ingest_probe.py
  run filter 120 python -u scripts/gpu_check.py filter
  run sort 120 python -u scripts/gpu_check.py sort
  run cub 120 scripts/ubench/build/cub_baseline
}

dist() { # N
  TAILN=25
  local n=${1:-2}
  run dist_check 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check.py
  run pytest_dist 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider
  TAILN=3
  run bench_n 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus "$n"
}

sanitize() { run sanitize 1700 bash scripts/sanitize.sh; }

"$@"
