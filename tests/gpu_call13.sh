#!/bin/bash
# fp16 running value in the stage-1 resblocks: parity (16-bit modes + everything touching the generator), role profile, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report13.jsonl
rm -f $TTSB_PARITY_REPORT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest13.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest13.log
tail -6 gpurun_out/pytest13.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles13.log
grep "umma_wa" gpurun_out/roles13.log | tail -24 | awk '{print $3,$4,$5,$6,$10,$11,$NF,$(NF-1)}' | sort | uniq -c
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops 2> gpurun_out/bench13.err | tee gpurun_out/bench13.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['profile_ms_per_step']['conv_tensor'], d['clocks']['sm_mhz'], d['workspace'])"
python tests/make_parity_table.py gpurun_out/parity_report13.jsonl | grep -E "f16|bf16" | cut -c1-330
