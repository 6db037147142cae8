#!/bin/bash
# round 2: Cantor prefix of the generic FFT schedule (GF(2^20): 12 passes instead of 17): parity + rate
tag=${1:-r2o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "additive or large_fields or gf2_20 or stages or golden or multiplicative or even_designed or closed_form" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
for v in "X=1" "BCHFFT_NO_CANTOR_PREFIX=1"; do
env $v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --batch 65536 --fft-batch 0 --only fft20 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - gpurun_out/${tag}_bench.json "$v" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
f=d["configs"]["additive_fft_gf2_20"]; sh=f["roofline"]["kernel_share_of_step"]; ms=f["ms_per_step"]
print("[%s] fft20 %.2f ms %.3e pts/s" % (sys.argv[2], ms, f["value"]), {k: round(ms*v,2) for k,v in sh.items()})
PY
done
