#!/bin/bash
# round 2, GPU call 1: the whole GPU suite (with the parity report), then the bench in the f16 and the strict fp32 tier
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
TTSB_CONV_LOG=1 timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_fp32.json | head -c 600
