#!/bin/bash
# wa planner experiments: fewer weight stages / smaller ring slots on the fp32-input layers with few taps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" TTSB_NO_PDL=1 TTSB_UMMA_PROF=1 timeout 300 python tests/profile_step.py f16 > /dev/null 2> gpurun_out/roles11_$tag.log
  echo "== $tag: $@"
  grep "umma_wa" gpurun_out/roles11_$tag.log | tail -24 | grep "xh=0" | awk '{print $5, $6, $10, $11, $NF, $(NF-1)}' | sort | uniq -c | head -12
  env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-hbm-ops 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['profile_ms_per_step']['conv_tensor'], d['clocks']['sm_mhz'])"
}
run base A=1
run ns2k3 TTSB_WA_NSTAGE=2 TTSB_WA_NSTAGE_KMAX=3
run ns2k7 TTSB_WA_NSTAGE=2 TTSB_WA_NSTAGE_KMAX=7
run rps1 TTSB_WA_RPS=1
run ns2k7rps1 TTSB_WA_NSTAGE=2 TTSB_WA_NSTAGE_KMAX=7 TTSB_WA_RPS=1
