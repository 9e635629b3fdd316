#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TTSB_LIB=$PWD/illufly-tts_b200/lib_prof/libttsb200.so TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles17.log
grep -A1 "umma pro=3 fmt=1 Cin=256 Cout=256 k=7" gpurun_out/roles17.log | tail -12 | cut -c1-330
