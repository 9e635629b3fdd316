#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python tests/bench_service.py 512 > gpurun_out/r02_service.md 2> gpurun_out/service33.err
echo "exit $?"; cat gpurun_out/r02_service.md | tail -12
