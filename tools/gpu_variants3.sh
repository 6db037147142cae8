#!/bin/bash
for v in "$@"; do
  make -C bch_fft_b200/csrc clean > /dev/null
  make -C bch_fft_b200/csrc -j16 all EXTRA_NVFLAGS="-DBCHFFT_M_LIST\(X\)=X\(13\)X\(16\) $v" > /dev/null 2> gpurun_out/build_err.log || { echo "build failed for $v"; tail -5 gpurun_out/build_err.log; continue; }
  python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --batch 32768 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
f=d["fft"]; sh=f["roofline"]["kernel_share_of_step"]; ms=f["ms_per_step"]
print("VARIANT [%s]: fft16 %.2f ms (ty %.2f gm %.2f)" % (sys.argv[1], ms, ms*sh.get("k_ty_pass",0), ms*sh["k_gm_pass"]))
PY
done
