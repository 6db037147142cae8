#!/bin/bash
# round 2, Berlekamp-Massey lane groups for large t: parity subset, then k_bm time per group size
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "at_scale or stages_vs_reference or wide_berlekamp or stages_and_decode" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in "X=1" "BCHFFT_BM_LPC=4" "BCHFFT_BM_LPC=8" "BCHFFT_BM_LPC=16" "BCHFFT_BM_LPC=32" "BCHFFT_BM_SM=0"; do
  env $v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --batch 131072 --fft-batch 0 --only t64,t200 > gpurun_out/${tag}_var.json 2> gpurun_out/${tag}_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/${tag}_var.err; continue; }
  python - "$v" gpurun_out/${tag}_var.json <<'PY'
import json, sys
d=json.load(open(sys.argv[2]))
for k in ("decode_n8191_t64", "decode_n8191_t200"):
    c=d["configs"][k]; ms=c["ms_per_step"]; sh=c["roofline"]["kernel_share_of_step"]
    print("VARIANT [%s] %s: %.2f ms %.3e cw/s" % (sys.argv[1], k, ms, c["value"]), {a: round(ms*b,2) for a,b in sh.items()}, c.get("parity"))
PY
done
