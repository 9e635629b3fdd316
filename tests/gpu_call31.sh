#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "conv1d or vocoder or config0 or config1 or batch_invariant or strict" 2>&1 | tail -4
for v in 0 1 0 1; do
if [ $v = 1 ]; then export TTSB_UMMA_NORESIDENT=1; else unset TTSB_UMMA_NORESIDENT; fi
timeout 200 python bench.py --steps 60 --warmup 5 --no-gpu-eager --no-hbm-ops --no-cpu-baseline > gpurun_out/bench31.json 2> gpurun_out/bench31.err; python -c "
import json; d=json.load(open('gpurun_out/bench31.json')); print('noresident=$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"
done
