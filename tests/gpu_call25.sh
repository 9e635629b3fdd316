#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 200 python bench.py --steps 60 --warmup 5 --no-gpu-eager --no-hbm-ops --no-cpu-baseline > gpurun_out/bench25.json 2> gpurun_out/bench25.err; python -c "
import json; d=json.load(open('gpurun_out/bench25.json')); print('$1', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"; }
run base
TTSB_WA_RPS16=4 run rps4
run base
TTSB_WA_RPS16=4 run rps4
TTSB_WA_RPS16=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vocoder or tensor_core" 2>&1 | tail -3
