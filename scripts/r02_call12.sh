source scripts/gpu_round.sh true
export TAILN=4
run sort_lb4 120 python -u scripts/gpu_check.py sort
VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_lb2.so run sort_lb2 120 python -u scripts/gpu_check.py sort
VINUM_B200_LIB=vinum_b200/_C/libvinum_b200_lb8.so run sort_lb8 120 python -u scripts/gpu_check.py sort
export TAILN=10
run pytest_sort 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py tests/test_gpu_topk.py -m gpu -q --maxfail=20 -p no:cacheprovider -k "sort or topk"
run ncu_sort 300 ncu --set full --clock-control none -f --import-source on -k regex:sort_pass -s 10 -c 1 -o gpurun_out/r02_sort_pass_lb python scripts/prof_kernels.py sort
