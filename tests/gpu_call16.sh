#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "config1 or vocoder or conv1d or strict or tensor_core or config0" 2>&1 | tail -5
timeout 300 python bench.py --steps 100 --warmup 5 --no-gpu-eager --no-hbm-ops > gpurun_out/bench16.json 2> gpurun_out/bench16.err; python -c "
import json; d=json.load(open('gpurun_out/bench16.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['profile_ms_per_step']['conv_tensor'])"
TTSB_LIB=$PWD/illufly-tts_b200/lib_prof/libttsb200.so TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles16.log
grep -A1 "umma pro=3 fmt=1 Cin=256 Cout=256" gpurun_out/roles16.log | tail -24 | cut -c1-300
