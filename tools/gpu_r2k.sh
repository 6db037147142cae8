#!/bin/bash
# round 2: single-call latencies after the wide-bucket rule; GF(2^20) with half tiles, two CTAs per SM
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --batch 131072 --fft-batch 0 --only single > gpurun_out/${tag}_single.json 2> gpurun_out/${tag}_single.err; echo "single rc=$?"
python - gpurun_out/${tag}_single.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
for k,v in d["configs"]["single_call"].items(): print(k, round(v["value"],4), v["unit"], v.get("parity"))
PY
bash tools/gpu_buildvar_fft20.sh "|X=1" "-DBCHFFT_NT_PASS=256 -DBCHFFT_MINB_PASS=2|BCHFFT_TMAX=10" "-DBCHFFT_NT_PASS=384|X=1"
