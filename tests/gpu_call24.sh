#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 200 python bench.py --steps 60 --warmup 5 --no-gpu-eager --no-hbm-ops --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err; python -c "
import json; d=json.load(open('gpurun_out/bench24.json')); print('$1', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['profile_ms_per_step']['conv_tensor'],3))"; }
run base
TTSB_WA_NSTAGE16=4 run ns4_k7
TTSB_WA_NSTAGE16=4 TTSB_WA_NSTAGE16_KMIN=11 run ns4_k11
TTSB_WA_NSTAGE16=5 TTSB_WA_NSTAGE16_KMIN=11 run ns5_k11
run base
