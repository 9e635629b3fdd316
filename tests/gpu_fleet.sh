#!/bin/bash
# BASELINE configs[4] on N GPUs of one box: bench.py --workload fleet (strong scaling).  usage: tests/gpu_fleet.sh N
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload fleet --steps 3 --warmup 1 \
  > gpurun_out/fleet_n$N.json 2> gpurun_out/fleet_n$N.err
echo "exit $?"; tail -3 gpurun_out/fleet_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/fleet_n{sys.argv[1]}.json") if l.startswith("{")][-1])
    print("fleet N =", d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "fleet", d["fleet"]["per_rank_device_ms"], d["fleet"]["imbalance_max_over_mean"])
except Exception as e:
    print("failed", e)
PY
