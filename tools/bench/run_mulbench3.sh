#!/bin/bash
for cfg in "-DMB_VARIANT=3" "-DMB_VARIANT=4" "-DMB_VARIANT=6"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bch_fft_b200/csrc $cfg tools/bench/mulbench.cu -o /tmp/mulbench 2>/dev/null && /tmp/mulbench
done
