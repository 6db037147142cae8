#!/bin/bash
tag=${1:-loc}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_locate_u" -c 1 -o /tmp/ncu_$tag python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --batch 262144 --fft-batch 0 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_src_$tag.csv 2>/dev/null
