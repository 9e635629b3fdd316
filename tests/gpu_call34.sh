#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench34.err; echo "exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value_pinned_host_buffer"], d["clocks"], d["roofline"]["frac"], d["gpu_launches"], d["cpu_baseline"]["value"], d["gpu_eager_baseline"]["value"])
PY
