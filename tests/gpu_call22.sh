#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vocoder or tensor_core" 2>&1 | tail -8
timeout 120 python tests/profile_step.py f16 2>&1 | tail -3
TTSB_NO_WA0=1 timeout 120 python tests/profile_step.py f16 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "config1 or config0 or config3" 2>&1 | tail -5
timeout 200 python bench.py --steps 100 --warmup 5 --no-gpu-eager --no-hbm-ops > gpurun_out/bench22.json 2> gpurun_out/bench22.err; python -c "
import json; d=json.load(open('gpurun_out/bench22.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"
