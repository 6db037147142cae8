#!/bin/bash
# round 2: warp-staged encoder (parity + rate)
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "encode or round_trip or stream" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 131072 --fft-batch 0 --no-e2e --only encode > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
python - gpurun_out/${tag}_bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
e=d["configs"]["encode_n8191_t16"]
print("encode %.3f ms %.4e cw/s" % (e["ms_per_step"], e["value"]), e["roofline"]["frac"], e["parity"], e["cpu_baseline"]["value"])
PY
