#!/bin/bash
# N GPUs, one process: positions-mode e2e through the library on all devices (after the concurrent-hosts hint)
N=${1:-8}
mkdir -p gpurun_out
for ft in default 1 4; do
  if [ "$ft" = default ]; then extra=""; else extra="BCHFFT_FLIP_THREADS=$ft"; fi
  env $extra BCHFFT_DEVICES=$(seq -s, 0 $((N-1))) BCHFFT_D2H_MODE=positions python tools/e2e_multi.py --child $N positions-$ft 1000000 0 2>> gpurun_out/r2zz.err | tee -a gpurun_out/r2zz_e2e_${N}gpu.jsonl
done
