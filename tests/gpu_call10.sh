#!/bin/bash
# two-stream F0 / N branches: full GPU suite, A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report10.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest10.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest10.log
tail -5 gpurun_out/pytest10.log
for mode in two one; do
  if [ $mode = one ]; then export TTSB_ONE_STREAM=1; else unset TTSB_ONE_STREAM; fi
  timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench10_$mode.json 2> gpurun_out/bench10_$mode.err
  tail -2 gpurun_out/bench10_$mode.err
  python - $mode <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench10_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "bench f16:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e)
PY
done
