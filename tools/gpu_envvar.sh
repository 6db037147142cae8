#!/bin/bash
# usage: gpu_envvar.sh "ENV1=a ENV2=b" "ENV1=c" ...   (FFT + decode bench per env setting, no rebuild)
timeout 900 python -m pytest tests -m gpu -q -x -k "additive or fft" > gpurun_out/pytest_env.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_env.log
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --batch 131072 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
f=d["fft"]; sh=f["roofline"]["kernel_share_of_step"]; ms=f["ms_per_step"]
print("VARIANT [%s]: fft16 %.2f ms" % (sys.argv[1], ms), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
