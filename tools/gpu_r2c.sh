#!/bin/bash
# round 2, call 3 (2 GPUs): patch-mode + multi-device tests, zero-copy and copy microbenchmarks, one-process e2e
tag=${1:-r2c}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt; nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1; nproc >> gpurun_out/${tag}_smi.txt
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q -x -k "patch_mode or multi_gpu or at_scale or stages" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
./tools/bench/zcbench 1000000 8 > gpurun_out/${tag}_zcbench.jsonl 2>&1; cat gpurun_out/${tag}_zcbench.jsonl
./tools/bench/zcbench 1000000 16 >> gpurun_out/${tag}_zcbench.jsonl 2>&1
./tools/bench/copybench 1024 5 > gpurun_out/${tag}_copybench.jsonl 2>&1; cat gpurun_out/${tag}_copybench.jsonl
timeout 600 python tools/e2e_multi.py 1000000 2048 > gpurun_out/${tag}_e2e_multi.jsonl 2> gpurun_out/${tag}_e2e_multi.err; echo "e2e_multi rc=$?"; cat gpurun_out/${tag}_e2e_multi.jsonl; tail -3 gpurun_out/${tag}_e2e_multi.err
