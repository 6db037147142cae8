#!/bin/bash
# usage: gpu_buildvar_fft20.sh "nvcc flags|env assignments" ...   (M = 13, 16, 20 only; GF(2^20) FFT bench per variant)
for v in "$@"; do
  flags="${v%%|*}"; envs="${v#*|}"
  make -C bch_fft_b200/csrc clean > /dev/null
  make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\)X\(20\) $flags" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed for $v"; tail -5 gpurun_out/build_err.log; continue; }
  env $envs python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --batch 65536 --fft-batch 0 --only fft20 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
f=d["configs"]["additive_fft_gf2_20"]; sh=f["roofline"]["kernel_share_of_step"]; ms=f["ms_per_step"]
print("VARIANT [%s]: fft20 %.2f ms %.3e pts/s" % (sys.argv[1], ms, f["value"]), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
