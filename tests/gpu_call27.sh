#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bilstm or single_utterance or ragged or edge or forced" 2>&1 | tail -4
timeout 120 python tests/profile_step.py f16 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "config1 or config0 or config4" 2>&1 | tail -3
for i in 1 2; do timeout 200 python bench.py --steps 100 --warmup 5 --no-gpu-eager --no-hbm-ops --no-cpu-baseline > gpurun_out/bench27.json 2> gpurun_out/bench27.err; python -c "
import json; d=json.load(open('gpurun_out/bench27.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3), round(d['profile_ms_per_step']['lstm'],3))"; done
python tests/bench_latency.py 2>&1 | tail -2
