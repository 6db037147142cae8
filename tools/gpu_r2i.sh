#!/bin/bash
# round 2: closed-form roots (L <= 2), k_locate_u buckets K = 6, 7; then FFT launch-shape variants
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "at_scale or stages_vs_reference or wide_berlekamp or stages_and_decode or golden or root_queue or closed_form or even_designed" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in "X=1" "BCHFFT_LOCATE_DIRECT=0" "BCHFFT_LOCU_NOSKIP=1"; do
env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --fft-batch 0 --only t64,t200 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - gpurun_out/${tag}_bench.json "$v" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ms=d["ms_per_step"]; sh=d["roofline"]["kernel_share_of_step"]
print("[%s] headline 1e6: %.3f ms %.4e cw/s" % (sys.argv[2], ms, d["value"]), {a: round(ms*b,3) for a,b in sh.items()})
for k in ("decode_n8191_t64", "decode_n8191_t200"):
    c=d["configs"][k]; ms=c["ms_per_step"]; sh=c["roofline"]["kernel_share_of_step"]
    print("%s: %.2f ms %.3e cw/s" % (k, ms, c["value"]), {a: round(ms*b,2) for a,b in sh.items()}, c.get("parity"))
PY
done
bash tools/gpu_buildvar_fft.sh "-DBCHFFT_NT_PASS=256 -DBCHFFT_MINB_PASS=2|BCHFFT_TMAX=10" "-DBCHFFT_NT_PASS=384|X=1" "-DBCHFFT_R4_MAD2=1|X=1" "-DBCHFFT_R4_MAD2=1 -DBCHFFT_NT_PASS=384|X=1"
