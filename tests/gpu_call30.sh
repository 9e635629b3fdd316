#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_invariant or batch32_full or vocoder or two_conv or edge or single_utterance" 2>&1 | tail -4
python tests/bench_latency.py 2>&1 | tail -2
TTSB_WA_NOSPLIT=1 python tests/bench_latency.py 2>&1 | tail -2
timeout 200 python bench.py --steps 60 --warmup 5 --no-gpu-eager --no-hbm-ops --no-cpu-baseline > gpurun_out/bench30.json 2> gpurun_out/bench30.err; python -c "
import json; d=json.load(open('gpurun_out/bench30.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
