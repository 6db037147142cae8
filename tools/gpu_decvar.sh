#!/bin/bash
# usage: gpu_decvar.sh "ENV=a" ...  decode parity tests once, then a decode bench per env setting (no rebuild)
timeout 900 python -m pytest tests -m gpu -q -x -k "bch or decode or golden" > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_dec.log
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --fft-batch 256 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
sh=d["roofline"]["kernel_share_of_step"]; ms=d["ms_per_step"]
print("VARIANT [%s]: decode %.2f ms %.3e cw/s" % (sys.argv[1], ms, d["value"]), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
