#!/bin/bash
for cfg in "-DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=32" "-DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=0" "-DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=4" "-DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=8" "-DMB_VARIANT=1" "-DMB_VARIANT=2" "-DMB_M=16 -DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=32" "-DMB_M=16 -DMB_VARIANT=0 -DBCHFFT_ALU_ROWS=4" "-DMB_M=16 -DMB_VARIANT=1"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bch_fft_b200/csrc $cfg tools/bench/mulbench.cu -o /tmp/mulbench 2>/dev/null && /tmp/mulbench
  cuobjdump -sass /tmp/mulbench | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -5 | tr '\n' ';'; echo
done
