#!/bin/bash
# rolling-window epilogue: parity subset, role profile, bench; then the per-config table with baselines and the fleet at N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "tensor_core or vocoder or config1_batch32 or strict or config0" > gpurun_out/pytest5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest5.log
tail -4 gpurun_out/pytest5.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles5.out 2> gpurun_out/roles5.log
grep "umma_wa" gpurun_out/roles5.log | tail -24 | cut -c1-330
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench5.json 2> gpurun_out/bench5.err
tail -2 gpurun_out/bench5.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench5.json").read().strip().splitlines()[-1])
    print("bench f16:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value_pinned_host_buffer"], d["clocks"], d["profile_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
timeout 900 python tests/bench_configs.py f16 > gpurun_out/configs5.md 2> gpurun_out/configs5.err
tail -3 gpurun_out/configs5.err; cat gpurun_out/configs5.md | cut -c1-200
timeout 600 python bench.py --workload fleet --steps 2 --warmup 1 > gpurun_out/fleet5_n1.json 2> gpurun_out/fleet5_n1.err
tail -2 gpurun_out/fleet5_n1.err; cut -c1-700 gpurun_out/fleet5_n1.json
