#!/bin/bash
# old conv kernel: bias one chunk ahead, FP64 statistics adds spread over the warp - parity subset, role profile, bench (f16 and strict)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "tensor_core or vocoder or config1_batch32 or strict or config0 or conv1d or fleet or chunk" > gpurun_out/pytest9.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest9.log
tail -4 gpurun_out/pytest9.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles9.out 2> gpurun_out/roles9.log
python tests/summarize_profile.py roles gpurun_out/roles9.log gpurun_out/roles9.md "TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 python tests/profile_step.py f16"; head -32 gpurun_out/roles9.md | tail -22
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench9.json 2> gpurun_out/bench9.err
tail -2 gpurun_out/bench9.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench9.json").read().strip().splitlines()[-1])
    print("bench f16:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["workspace"], d["profile_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp32 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench9_fp32.json 2> gpurun_out/bench9_fp32.err
cut -c1-200 gpurun_out/bench9_fp32.json
