#!/bin/bash
# round 2, second GPU call: the new all-configs bench line + launch list + --set full captures
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 300 python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 128 > gpurun_out/${tag}_bench_small.json 2> gpurun_out/${tag}_bench_small.err; echo "small bench rc=$?"; tail -3 gpurun_out/${tag}_bench_small.err
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err
CMD="python bench.py --steps 2 --warmup 1 --batch 131072 --fft-batch 1024 --no-cpu-baseline --no-e2e --only fft"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ncu.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_syndrome_p|k_bm|k_elp_fwd_u|k_locate_u|k_correct" -c 5 -o /tmp/ncu_dec_$tag $CMD > gpurun_out/${tag}_ncu_full_dec.log 2>&1; echo "ncu decode rc=$?"
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_decode.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bitslice_in|k_tx_line_pass|k_ty_plane_pass|k_gm_pass_pm|k_gather_out" -c 6 -o /tmp/ncu_fft_$tag $CMD > gpurun_out/${tag}_ncu_full_fft.log 2>&1; echo "ncu fft rc=$?"
ncu -i /tmp/ncu_fft_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft.csv 2>/dev/null
# side configurations: t = 64 (general root search), GF(2^20), multiplicative
CMD2="python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 0 --no-cpu-baseline --no-e2e --only t64,fft20,mfft"
timeout 900 ncu --set full --clock-control none -k regex:"^k_locate$|^k_elp_fwd$|^k_gm_pass$|^k_ty_pass$|^k_mfft_stage$|^k_syndrome$" -c 36 -o /tmp/ncu_side_$tag $CMD2 > gpurun_out/${tag}_ncu_full_side.log 2>&1; echo "ncu side rc=$?"
ncu -i /tmp/ncu_side_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_side.csv 2>/dev/null
ls -la gpurun_out | grep $tag
