#!/bin/bash
# full GPU check: whole -m gpu suite, smoke, full bench (both arms); tag = output suffix
tag=${1:-full}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("decode cw/s %.3e  ms %.2f  e2e %.3e" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value",0)))
print(" shares", d["roofline"]["kernel_share_of_step"])
f=d["fft"]
if f: print("fft pts/s %.3e ms %.2f e2e %.3e" % (f["value"], f["ms_per_step"], (f["e2e"] or {}).get("value",0)), f["roofline"]["kernel_share_of_step"])
print(d.get("extras"))
PY
