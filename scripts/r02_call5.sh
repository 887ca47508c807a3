# 1 GPU: new tests, sanitizer, ingest probe, ncu captures, launch list
source scripts/gpu_round.sh true
export TAILN=8
run pytest_dropin 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_paths.py -m gpu -q --maxfail=10 -p no:cacheprovider
for th in 4 8 12; do VINUM_B200_INGEST_THREADS=$th run ingest_t$th 300 python -u scripts/ingest_probe.py; done
VINUM_B200_SORT_RANK=0 run sort_rank0 120 python -u scripts/gpu_check.py sort
VINUM_B200_SORT_RANK=1 run sort_rank1 120 python -u scripts/gpu_check.py sort
NCU="ncu --set full --clock-control none -f"
run ncu_filter 300 $NCU --import-source on -k regex:filter_kernel -s 2 -c 1 -o gpurun_out/r02_filter python scripts/prof_kernels.py filter
run ncu_northstar 400 $NCU --import-source on -k regex:agg_fast -s 3 -c 1 -o gpurun_out/r02_agg_fast python scripts/prof_kernels.py northstar
run ncu_hash 400 $NCU --import-source on -k regex:agg_fast -s 3 -c 1 -o gpurun_out/r02_agg_fast_dict python scripts/prof_kernels.py hash
run ncu_c3 400 $NCU --import-source on -k regex:agg_fast -s 3 -c 1 -o gpurun_out/r02_agg_fast_c3 python scripts/prof_kernels.py c3
run ncu_sort 300 $NCU --import-source on -k regex:sort_pass -s 10 -c 1 -o gpurun_out/r02_sort_pass python scripts/prof_kernels.py sort
run ncu_arith 200 $NCU -k regex:arith8 -s 2 -c 1 -o gpurun_out/r02_arith8 python scripts/prof_kernels.py arith
run ncu_compare 200 $NCU -k regex:compare8 -s 2 -c 1 -o gpurun_out/r02_compare8 python scripts/prof_kernels.py compare
run ncu_onegroup 200 $NCU -k regex:agg_onegroup8 -s 4 -c 1 -o gpurun_out/r02_onegroup8 python scripts/prof_kernels.py onegroup
run ncu_topk 200 $NCU -k regex:topk_hist -s 2 -c 1 -o gpurun_out/r02_topk_hist python scripts/prof_kernels.py topk
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-verify --no-configs
run sanitize 1500 bash scripts/sanitize.sh
ls -la gpurun_out/*.ncu-rep
