#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 4 0 4 5; do
TTSB_L2PF=$v timeout 300 python bench.py --steps 60 --warmup 5 --no-gpu-eager --no-hbm-ops > gpurun_out/bench18_$v.json 2> gpurun_out/bench18.err; python -c "
import json; d=json.load(open('gpurun_out/bench18_$v.json')); print('L2PF=$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"
done
