#!/bin/bash
# round 2: launch list + --set full captures (decode kernels with source view, FFT kernels) of the final code
tag=${1:-r2fin}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --batch 131072 --fft-batch 1024 --no-cpu-baseline --no-e2e --only fft"
timeout 600 $CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ncu.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_syndrome_p|k_bm|k_bucket|k_elp_fwd_u|k_locate_u|k_correct" -c 7 -o /tmp/ncu_dec_$tag $CMD > gpurun_out/${tag}_ncu_full_dec.log 2>&1; echo "ncu decode rc=$?"
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_decode.csv 2>/dev/null
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page source --csv -k regex:k_locate_u > gpurun_out/${tag}_ncu_src_locate.csv 2>/dev/null
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page source --csv -k regex:k_syndrome_p > gpurun_out/${tag}_ncu_src_syndrome.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_bitslice_in|k_tx_line_pass|k_ty_plane_pass|k_gm_pass_pm|k_gather_out" -c 6 -o /tmp/ncu_fft_$tag $CMD > gpurun_out/${tag}_ncu_full_fft.log 2>&1; echo "ncu fft rc=$?"
ncu -i /tmp/ncu_fft_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft.csv 2>/dev/null
ls -la gpurun_out | grep $tag
