#!/bin/bash
# round 2, first GPU call: whole GPU suite (no -x), copy ceiling, sanitizer passes, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=25 ) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
./tools/bench/copybench 1024 5 > gpurun_out/r2a_copybench_1gpu.jsonl 2>&1; cat gpurun_out/r2a_copybench_1gpu.jsonl
tools/gpu_sanitize.sh r2a
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
