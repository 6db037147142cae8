#!/bin/bash
# build and time kernel variants on the GPU box: each argument is a set of extra nvcc flags
for v in "$@"; do
  make -C bch_fft_b200/csrc clean > /dev/null
  make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\) $v" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed for $v"; tail -5 gpurun_out/build_err.log; continue; }
  grep -h -A2 -E "k_gm_passILi16|k_locateILi13" bch_fft_b200/csrc/build/*.ptxas.log | grep -E "spill|Used" | tr '\n' ' '
  echo
  python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --fft-batch 2048 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
sh=d["roofline"]["kernel_share_of_step"]; ms=d["ms_per_step"]
f=d["fft"]
print("VARIANT [%s]: decode %.2f ms (locate %.2f syn %.2f bm %.2f) | fft16 %.2f ms" % (sys.argv[1], ms, ms*sh["k_locate"], ms*sh["k_syndrome"], ms*sh["k_bm"], f["ms_per_step"]))
PY
done
