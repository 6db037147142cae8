#!/bin/bash
for v in "$@"; do
  env $v python bench.py --steps 3 --warmup 2 --no-cpu-baseline --fft-batch 0 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/bench_var.err; continue; }
  python - "$v" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_var.json"))
print("VARIANT [%s]: decode %.2f ms  e2e %.3e cw/s" % (sys.argv[1], d["ms_per_step"], d["e2e"]["value"]))
PY
done
