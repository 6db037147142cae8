#!/bin/bash
for m in 13 16; do for cfg in "-DMB_VARIANT=4" "-DMB_VARIANT=7"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bch_fft_b200/csrc -DMB_M=$m $cfg tools/bench/mulbench.cu -o /tmp/mulbench 2>/dev/null && /tmp/mulbench
done; done
