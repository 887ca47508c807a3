# one gpurun call of round 2: parity suite, kernel probes, bench
source scripts/gpu_round.sh true
rm -f gpurun_out/round.log
export TAILN=${TAILN:-6}
"$@"
