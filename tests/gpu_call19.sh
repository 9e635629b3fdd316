#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tests/profile_latency.py 2>&1 | tee gpurun_out/latency_profile19.md
TTSB_NO_PDL=1 timeout 200 python tests/profile_latency.py 2>&1 | head -1
