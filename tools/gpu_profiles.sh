#!/bin/bash
# round profiles: launch list of one bench command + one `--set full` capture of the hot kernels
tag=${1:-r1h}
CMD="python bench.py --steps 2 --warmup 2 --batch 131072 --fft-batch 1024 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_syndrome_p|k_bm|k_elp_fwd_u|k_locate_u" -c 4 -o /tmp/ncu_dec_$tag $CMD > gpurun_out/ncu_full_dec_$tag.log 2>&1; echo "ncu decode rc=$?"
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_decode.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bitslice_in|k_tx_line_pass|k_ty_plane_pass|k_gm_pass_pm|k_gather_out" -c 6 -o /tmp/ncu_fft_$tag $CMD > gpurun_out/ncu_full_fft_$tag.log 2>&1; echo "ncu fft rc=$?"
ncu -i /tmp/ncu_fft_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft.csv 2>/dev/null
ls -la gpurun_out | head -30
