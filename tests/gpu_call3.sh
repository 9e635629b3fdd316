#!/bin/bash
# wa kernel v2 + workspace scopes: the whole GPU suite, role profile, bench (with the GPU-eager baseline)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report3.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest3.log
tail -6 gpurun_out/pytest3.log
TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles3.out 2> gpurun_out/roles3.log
grep "umma_wa" gpurun_out/roles3.log | tail -24 | cut -c1-330
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_f16.json 2> gpurun_out/bench3_f16.err
tail -3 gpurun_out/bench3_f16.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench3_f16.json").read().strip().splitlines()[-1])
    print("bench f16:", d["value"], d["ms_per_step"], d["e2e"], d["profile_ms_per_step"], d.get("gpu_eager_baseline"))
except Exception as e:
    print("bench failed", e)
PY
