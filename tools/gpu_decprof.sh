#!/bin/bash
tag=${1:-dec}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_locate_u|k_syndrome|k_bm" -c 3 -o /tmp/ncu_$tag python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --batch 262144 --fft-batch 64 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv -k regex:k_locate_u > gpurun_out/ncu_src_locate_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page details -k regex:k_locate_u > gpurun_out/ncu_details_locate_$tag.txt 2>/dev/null
ls -la /tmp/ncu_$tag.ncu-rep
