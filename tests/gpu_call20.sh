#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 TTSB_WA_PROF=1 timeout 200 python tests/bench_latency.py > /dev/null 2> gpurun_out/roles20.log
wc -l gpurun_out/roles20.log
