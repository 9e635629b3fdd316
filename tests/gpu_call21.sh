#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-gpu-eager --no-hbm-ops > gpurun_out/bench21_$i.json 2> gpurun_out/bench21.err; python -c "
import json; d=json.load(open('gpurun_out/bench21_$i.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"
done
