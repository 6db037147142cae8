#!/bin/bash
# usage: gpu_buildvar_fft.sh "nvcc flags|env assignments" ...   (M = 13, 16 only; FFT bench per variant)
for v in "$@"; do
  flags="${v%%|*}"; envs="${v#*|}"
  make -C bch_fft_b200/csrc clean > /dev/null
  make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\) $flags" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed for $v"; tail -5 gpurun_out/build_err.log; continue; }
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_additive_fft_vs_oracle and (13 or 16)" > gpurun_out/pytest_var.log 2>&1; tail -1 gpurun_out/pytest_var.log
  env $envs python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --skip-configs --batch 65536 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
f=d["fft"]; sh=f["roofline"]["kernel_share_of_step"]; ms=f["ms_per_step"]
print("VARIANT [%s]: fft16 %.2f ms %.3e pts/s" % (sys.argv[1], ms, f["value"]), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
