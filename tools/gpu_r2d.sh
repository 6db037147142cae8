#!/bin/bash
# round 2, call 4: new multiplicative FFT (GF(2^16) fast path) + positions-mode e2e
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "multiplicative or patch_mode or bchstream or concurrent or full_batch_properties" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --only mfft > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("decode cw/s %.3e ms %.2f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
m=d["configs"]["multiplicative_fft_gf2_16"]
print("mfft pts/s %.3e ms %.2f e2e %.3e" % (m["value"], m["ms_per_step"], m["e2e"]["value"]), m["roofline"]["kernel_share_of_step"], m["parity"])
PY
for mode in positions words; do for ft in 2 4 8; do BCHFFT_FLIP_THREADS=$ft BCHFFT_D2H_MODE=$mode python tools/e2e_multi.py --child 1 $mode-$ft 1000000 0 >> gpurun_out/${tag}_e2e_modes.jsonl 2>> gpurun_out/${tag}_e2e_modes.err; done; done; cat gpurun_out/${tag}_e2e_modes.jsonl
CMD="python bench.py --small --steps 2 --warmup 1 --batch 32768 --fft-batch 0 --no-cpu-baseline --no-e2e --only mfft"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mfft16" -c 4 -o /tmp/ncu_mfft_$tag $CMD > gpurun_out/${tag}_ncu_full_mfft.log 2>&1; echo "ncu mfft rc=$?"
ncu -i /tmp/ncu_mfft_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_mfft.csv 2>/dev/null
cp /tmp/ncu_mfft_$tag.ncu-rep gpurun_out/${tag}_mfft.ncu-rep
