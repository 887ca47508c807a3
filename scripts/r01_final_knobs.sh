# Full GPU parity suite with the defaults as shipped (the knobs that won the first-contact probe are
# now the defaults), the top-k tests in a second process to fit the round's last GPU seconds.
mkdir -p gpurun_out
(timeout 30 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_topk.py > gpurun_out/final_defaults_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/final_defaults_suite.log) &
(timeout 30 python -m pytest tests/test_gpu_topk.py -m gpu -x -q > gpurun_out/final_defaults_topk.log 2>&1; echo "topk rc=$?" >> gpurun_out/final_defaults_topk.log) &
wait
tail -3 gpurun_out/final_defaults_suite.log
tail -3 gpurun_out/final_defaults_topk.log | cut -c1-400
