#!/bin/bash
# round 2: ncu capture of the GF(2^20) pass kernel only
tag=${1:-r2s}
mkdir -p gpurun_out
CMD3="python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 0 --no-cpu-baseline --no-e2e --only fft20"
timeout 900 ncu --set full --clock-control none -k regex:"^k_gm_pass$" -c 12 -o /tmp/ncu_fft20_$tag $CMD3 > gpurun_out/${tag}_ncu_full_fft20.log 2>&1; echo "ncu fft20 rc=$?"
ncu -i /tmp/ncu_fft20_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft20.csv 2>/dev/null
ls -la gpurun_out | grep $tag
