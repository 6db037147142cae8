#!/bin/bash
# FFT-focused: parity tests for the additive path, short bench, ncu full of the FFT kernels
tag=${1:-fft}
timeout 900 python -m pytest tests -m gpu -q -x -k "additive or fft" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("decode cw/s %.3e  ms %.2f" % (d["value"], d["ms_per_step"]))
f=d["fft"]
ms=f["ms_per_step"]
print("fft pts/s %.3e ms %.2f" % (f["value"], ms), {k: round(ms*v,3) for k,v in f["roofline"]["kernel_share_of_step"].items()})
print(d.get("extras"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ty_plane_pass|k_gm_pass_pm|k_bitslice_in|k_gather_out" -c 6 -o /tmp/ncu_$tag python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --batch 32768 --fft-batch 1024 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv -k regex:k_ty_plane_pass > gpurun_out/ncu_src_plane_$tag.csv 2>/dev/null
ls -la /tmp/ncu_$tag.ncu-rep gpurun_out/ | head -5
