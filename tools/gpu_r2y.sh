#!/bin/bash
# k_locate_u root-queue capacity (shared memory not used by the queue is L1 for the forward-value loads)
mkdir -p gpurun_out
for v in "BCHFFT_LOCU_QUEUE=4096" "BCHFFT_LOCU_QUEUE=2048" "BCHFFT_LOCU_QUEUE=1024" "BCHFFT_LOCU_QUEUE=512" "BCHFFT_LOCU_QUEUE=256" "BCHFFT_LOCU_QUEUE=128"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --fft-batch 0 --only t64 > gpurun_out/r2y_var.json 2> gpurun_out/r2y_var.err || { echo "bench failed for $v"; tail -3 gpurun_out/r2y_var.err; continue; }
  python - "$v" gpurun_out/r2y_var.json <<'PY'
import json, sys
d=json.load(open(sys.argv[2]))
ms=d["ms_per_step"]; sh=d["roofline"]["kernel_share_of_step"]
c=d["configs"]["decode_n8191_t64"]; ms2=c["ms_per_step"]; sh2=c["roofline"]["kernel_share_of_step"]
print("VARIANT [%s] t16 %.3f ms locate %.3f | t64 %.3f ms locate %.3f" % (sys.argv[1], ms, ms*sh["k_locate_u"], ms2, ms2*sh2["k_locate_u"]))
PY
done
