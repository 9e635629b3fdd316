#!/bin/bash
# table entries prefetched ahead (ItemWalker) in every role of both conv kernels: parity, role profile, bench, latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "not fp32 or strict or single_utterance or ragged or edge" > gpurun_out/pytest14.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest14.log
tail -4 gpurun_out/pytest14.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles14.log
python tests/summarize_profile.py roles gpurun_out/roles14.log gpurun_out/roles14.md "TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 python tests/profile_step.py f16"; head -24 gpurun_out/roles14.md | tail -14
grep "umma_wa" gpurun_out/roles14.log | tail -24 | awk '{s+=$(NF-1)} END{print "wa sum us (prof mode)", s}'
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops 2> gpurun_out/bench14.err | tee gpurun_out/bench14.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['profile_ms_per_step']['conv_tensor'], d['clocks']['sm_mhz'], d['workspace']['peak_bytes'])"
python tests/bench_latency.py 2>&1 | tail -3
