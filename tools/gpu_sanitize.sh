#!/bin/bash
# compute-sanitizer passes over tools/sanitize_workload.py; logs -> gpurun_out/sanitize_*.log
# usage: tools/gpu_sanitize.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, level, extra env
  local tool=$1 level=$2; shift 2
  env "$@" timeout ${SAN_TIMEOUT:-420} $CS --tool $tool --print-limit 30 --error-exitcode 9 \
      python tools/sanitize_workload.py $level > gpurun_out/sanitize_${tag}_${tool}${SUFFIX}.log 2>&1
  echo "$tool $level $* -> exit $?" | tee -a gpurun_out/sanitize_${tag}_summary.txt
  tail -3 gpurun_out/sanitize_${tag}_${tool}${SUFFIX}.log | tee -a gpurun_out/sanitize_${tag}_summary.txt
}
run memcheck full X=1
SUFFIX=_queue6 run memcheck light BCHFFT_LOCU_QUEUE=6
run synccheck full X=1
run racecheck light X=1
SUFFIX=_queue6 run racecheck light BCHFFT_LOCU_QUEUE=6
run initcheck light X=1
