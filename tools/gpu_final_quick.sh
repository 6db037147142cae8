#!/bin/bash
# last sanity check of the final build: a parity subset, smoke, headline + FFT bench with e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or concurrent or stream or patch_mode or multi_gpu or routes or native" > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --skip-configs > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("decode %.4e cw/s (%.3f ms) e2e %.4e parity %s; fft %.4e e2e %.4e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"], d["fft"]["value"], d["fft"]["e2e"]["value"]))
PY
