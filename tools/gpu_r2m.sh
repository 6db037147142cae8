#!/bin/bash
# round 2 (N GPUs): in-library sharding test, one-process e2e over 1..N devices (positions / words), torchrun bench at N
tag=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt; nproc >> gpurun_out/${tag}_smi.txt
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q -x -k "multi_gpu" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 900 python tools/e2e_multi.py 1000000 ${3:-2048} > gpurun_out/${tag}_e2e_one_process.jsonl 2> gpurun_out/${tag}_e2e_one_process.err; echo "e2e_multi rc=$?"; cat gpurun_out/${tag}_e2e_one_process.jsonl; tail -3 gpurun_out/${tag}_e2e_one_process.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --skip-configs > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "torchrun bench rc=$?"; tail -2 gpurun_out/${tag}_bench_${N}gpu.err
python - gpurun_out/${tag}_bench_${N}gpu.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d decode %.4e cw/s (%.3f ms) e2e %.4e; fft %.4e pts/s e2e %.4e" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["fft"]["value"], d["fft"]["e2e"]["value"]))
PY
