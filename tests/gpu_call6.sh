#!/bin/bash
# old conv kernel: residual loads at the chunk top, elected arrivals, L2 prefetch one item ahead - parity subset, role profile, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "tensor_core or vocoder or config1_batch32 or strict or config0 or conv1d or fleet or chunk" > gpurun_out/pytest6.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest6.log
tail -4 gpurun_out/pytest6.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles6.out 2> gpurun_out/roles6.log
python tests/summarize_profile.py roles gpurun_out/roles6.log gpurun_out/roles6.md "TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 python tests/profile_step.py f16"; head -45 gpurun_out/roles6.md | tail -36
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -2 gpurun_out/bench6.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench6.json").read().strip().splitlines()[-1])
    print("bench f16:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value_pinned_host_buffer"], d["clocks"], d["profile_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp32 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench6_fp32.json 2> gpurun_out/bench6_fp32.err
cut -c1-200 gpurun_out/bench6_fp32.json
