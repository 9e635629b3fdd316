#!/bin/bash
# the part of tests/gpu_final.sh that depends on the engine binary's frame side (no reference arm, no strict-tier bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report_final.jsonl
rm -f $TTSB_PARITY_REPORT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log; tail -4 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_final.log
python tests/bench_latency.py > gpurun_out/r02_latency.md 2>&1; cat gpurun_out/r02_latency.md
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; tail -2 gpurun_out/bench_final.err
timeout 900 python tests/bench_configs.py f16 > gpurun_out/r02_configs.md 2> gpurun_out/configs_final.err; echo "configs exit $?"
python - <<'PY'
import json
for f in ("r02_bench_final",):
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d.get("e2e"), d.get("clocks"), d.get("cpu_baseline", {}).get("value"), (d.get("gpu_eager_baseline") or {}).get("value"))
PY
