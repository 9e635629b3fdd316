#!/bin/bash
# A/B of the L2 prefetch: bit 0 = old kernel, bit 1 = weights-as-A kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 2 1 3; do
  TTSB_L2PF=$v timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench7_$v.json 2> gpurun_out/bench7_$v.err
  python - $v <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench7_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("L2PF", sys.argv[1], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["profile_ms_per_step"]["conv_tensor"])
except Exception as e:
    print("bench failed", e)
PY
done
