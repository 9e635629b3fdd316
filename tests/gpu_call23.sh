#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles23.log
grep "umma_wa C=256" gpurun_out/roles23.log | tail -24 | cut -c1-330
