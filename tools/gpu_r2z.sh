#!/bin/bash
# final check of the round: k_locate_u CTA target on one box (2400 = the old default), then the whole
# -m gpu suite, smoke and both bench arms of the code as committed
mkdir -p gpurun_out
for v in "BCHFFT_LOCU_CTAS=2400" "BCHFFT_LOCU_CTAS=16000"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --fft-batch 0 --only t64 > gpurun_out/r2z_var.json 2> gpurun_out/r2z_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/r2z_var.err; continue; }
  python - "$v" gpurun_out/r2z_var.json <<'PY' | tee -a gpurun_out/r2z_locu_ctas.txt
import json, sys
d=json.load(open(sys.argv[2]))
ms=d["ms_per_step"]; sh=d["roofline"]["kernel_share_of_step"]
c=d["configs"]["decode_n8191_t64"]; ms2=c["ms_per_step"]; sh2=c["roofline"]["kernel_share_of_step"]
print("VARIANT [%s] t16 %.3f ms locate %.3f | t64 %.3f ms locate %.3f" % (sys.argv[1], ms, ms*sh["k_locate_u"], ms2, ms2*sh2["k_locate_u"]))
PY
done
bash tools/gpu_r2v.sh r2z
