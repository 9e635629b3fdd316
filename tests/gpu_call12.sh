#!/bin/bash
# text encoder on the second stream, k3 two weight stages: parity subset, bench A/B, latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "not fp32 or strict or single_utterance or ragged" > gpurun_out/pytest12.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest12.log
tail -4 gpurun_out/pytest12.log
for mode in two one; do
  if [ $mode = one ]; then export TTSB_ONE_STREAM=1; else unset TTSB_ONE_STREAM; fi
  timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops 2> gpurun_out/bench12_$mode.err | tee gpurun_out/bench12_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['profile_ms_per_step']['conv_tensor'], d['clocks']['sm_mhz'])"
done
unset TTSB_ONE_STREAM
python tests/bench_latency.py 2>&1 | tail -4
