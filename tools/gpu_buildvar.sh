#!/bin/bash
# usage: gpu_buildvar.sh "nvcc flags|env assignments" ...   (M = 13, 16 only; decode bench per variant)
for v in "$@"; do
  flags="${v%%|*}"; envs="${v#*|}"
  make -C bch_fft_b200/csrc clean > /dev/null
  make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\) $flags" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed for $v"; tail -5 gpurun_out/build_err.log; continue; }
  env $envs python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --skip-configs --fft-batch 0 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
sh=d["roofline"]["kernel_share_of_step"]; ms=d["ms_per_step"]
print("VARIANT [%s]: decode %.2f ms %.3e cw/s" % (sys.argv[1], ms, d["value"]), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
