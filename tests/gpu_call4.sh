#!/bin/bash
# wa epilogue v3 + programmatic dependent launch: parity subset, role profile, bench with and without PDL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "tensor_core or vocoder or config1_batch32 or strict or config0" > gpurun_out/pytest4.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest4.log
tail -4 gpurun_out/pytest4.log
TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > gpurun_out/roles4.out 2> gpurun_out/roles4.log
grep "umma_wa" gpurun_out/roles4.log | tail -24 | cut -c1-330
for mode in pdl nopdl; do
  if [ $mode = nopdl ]; then export TTSB_NO_PDL=1; else unset TTSB_NO_PDL; fi
  timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops > gpurun_out/bench4_$mode.json 2> gpurun_out/bench4_$mode.err
  tail -2 gpurun_out/bench4_$mode.err
  python - $mode <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench4_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "bench f16:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value_pinned_host_buffer"], d["profile_ms_per_step"])
except Exception as e:
    print("bench failed", e)
PY
done
unset TTSB_NO_PDL
python tests/bench_latency.py > gpurun_out/latency4.txt 2>&1; tail -8 gpurun_out/latency4.txt
