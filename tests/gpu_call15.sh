#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles15.log
grep -A1 "umma pro=3 fmt=1 Cin=256 Cout=256" gpurun_out/roles15.log | tail -24 | cut -c1-300
grep -A1 "Cin=128 Cout=22\|Cin=22 Cout=128\|Cin=1090 Cout=1024 k=3\|Cin=768 Cout=2048" gpurun_out/roles15.log | tail -12 | cut -c1-300
