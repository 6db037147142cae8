#!/bin/bash
tag=${1:-bf}
make -C bch_fft_b200/csrc clean > /dev/null
make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\)" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed"; tail -5 gpurun_out/build_err.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gm_pass_pm" -c 2 -o /tmp/ncu_$tag python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --batch 32768 --fft-batch 1024 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_src_$tag.csv 2>/dev/null
