#!/bin/bash
# round 2: new k_syndrome_p + Berlekamp-Massey lane groups: parity subset, variants, multiply microbenchmark
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "at_scale or stages_vs_reference or wide_berlekamp or stages_and_decode or golden or full_batch_properties" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in "X=1" "BCHFFT_BM_LPC=2" "BCHFFT_BM_LPC=4" "BCHFFT_BM_LPC=8" "BCHFFT_BM_LPC=16" "BCHFFT_BM_LPC=32"; do
  env $v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --batch 131072 --fft-batch 0 --only t64,t200 > gpurun_out/${tag}_var.json 2> gpurun_out/${tag}_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/${tag}_var.err; continue; }
  python - "$v" gpurun_out/${tag}_var.json <<'PY'
import json, sys
d=json.load(open(sys.argv[2]))
ms=d["ms_per_step"]; sh=d["roofline"]["kernel_share_of_step"]
print("VARIANT [%s] headline(131072): %.3f ms" % (sys.argv[1], ms), {a: round(ms*b,3) for a,b in sh.items()})
for k in ("decode_n8191_t64", "decode_n8191_t200"):
    c=d["configs"][k]; ms=c["ms_per_step"]; sh=c["roofline"]["kernel_share_of_step"]
    print("VARIANT [%s] %s: %.2f ms %.3e cw/s" % (sys.argv[1], k, ms, c["value"]), {a: round(ms*b,2) for a,b in sh.items()}, c.get("parity"))
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --fft-batch 0 --skip-configs > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - gpurun_out/${tag}_bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ms=d["ms_per_step"]; sh=d["roofline"]["kernel_share_of_step"]
print("headline 1e6: %.3f ms %.4e cw/s" % (ms, d["value"]), {a: round(ms*b,3) for a,b in sh.items()})
PY
bash tools/bench/run_mulbench4.sh
