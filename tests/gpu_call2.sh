#!/bin/bash
# round 2, GPU call 2: the weights-as-A kernel on stage 1 - parity of the 16-bit modes, bench, role profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report2.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "f16 or bf16 or tensor or vocoder or strict or batch32" > gpurun_out/pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest2.log
tail -4 gpurun_out/pytest2.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_f16.json 2> gpurun_out/bench2_f16.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench2_f16.json").read().strip().splitlines()[-1])
    print("bench f16:", d["value"], d["ms_per_step"], d["profile_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles2.out 2> gpurun_out/roles2.log
grep -c umma gpurun_out/roles2.log
