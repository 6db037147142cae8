#!/bin/bash
# round 2, full check of the final code: whole -m gpu suite, smoke, both bench arms, launch list, --set full captures
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/${tag}_bench.err
CMD="python bench.py --steps 2 --warmup 1 --batch 131072 --fft-batch 1024 --no-cpu-baseline --no-e2e --only fft"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ncu.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_syndrome_p|k_bm|k_bucket|k_elp_fwd_u|k_locate_u|k_correct" -c 7 -o /tmp/ncu_dec_$tag $CMD > gpurun_out/${tag}_ncu_full_dec.log 2>&1; echo "ncu decode rc=$?"
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_decode.csv 2>/dev/null
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page source --csv -k regex:k_locate_u > gpurun_out/${tag}_ncu_src_locate.csv 2>/dev/null
ncu -i /tmp/ncu_dec_$tag.ncu-rep --page source --csv -k regex:k_syndrome_p > gpurun_out/${tag}_ncu_src_syndrome.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bitslice_in|k_tx_line_pass|k_ty_plane_pass|k_gm_pass_pm|k_gather_out" -c 6 -o /tmp/ncu_fft_$tag $CMD > gpurun_out/${tag}_ncu_full_fft.log 2>&1; echo "ncu fft rc=$?"
ncu -i /tmp/ncu_fft_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft.csv 2>/dev/null
# side configurations: t = 64 / 200 (lane-group Berlekamp-Massey, wide buckets), GF(2^20), multiplicative
CMD2="python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 0 --no-cpu-baseline --no-e2e --only t64,t200,fft20,mfft"
timeout 900 ncu --set full --clock-control none -k regex:"^k_locate$|^k_locate_u$|^k_elp_fwd$|^k_bm$|^k_gm_pass$|^k_ty_pass$|k_mfft16|^k_syndrome_p$" -c 60 -o /tmp/ncu_side_$tag $CMD2 > gpurun_out/${tag}_ncu_full_side.log 2>&1; echo "ncu side rc=$?"
ncu -i /tmp/ncu_side_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_side.csv 2>/dev/null
ls -la gpurun_out | grep $tag
# GF(2^20) generic schedule: its pass kernel on its own (the side capture above stops before it)
CMD3="python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 0 --no-cpu-baseline --no-e2e --only fft20"
timeout 900 ncu --set full --clock-control none -k regex:"^k_gm_pass$" -c 12 -o /tmp/ncu_fft20_$tag $CMD3 > gpurun_out/${tag}_ncu_full_fft20.log 2>&1; echo "ncu fft20 rc=$?"
ncu -i /tmp/ncu_fft20_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_fft20.csv 2>/dev/null
