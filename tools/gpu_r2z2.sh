#!/bin/bash
# launch list + --set full capture of k_locate_u (with source view) for the code with the 16000-CTA target,
# at the bench's full batch (10^6 words) so that grid and duration are the ones of the bench line
tag=${1:-r2z}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --fft-batch 1024 --no-cpu-baseline --no-e2e --only fft"
timeout 300 $CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ncu.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_locate_u" -c 1 -o /tmp/ncu_dec_$tag $CMD > gpurun_out/${tag}_ncu_full_dec.log 2>&1; echo "ncu locate rc=$?"
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_locate.csv 2>/dev/null
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page source --csv -k regex:k_locate_u > gpurun_out/${tag}_ncu_src_locate.csv 2>/dev/null
ls -la gpurun_out | grep $tag
