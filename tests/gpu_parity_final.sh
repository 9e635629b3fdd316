#!/bin/bash
# parity report of the final binary: tests/test_gpu_configs.py writes one JSON line per (config, mode)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TTSB_PARITY_REPORT=$PWD/gpurun_out/parity_report_final.jsonl
rm -f $TTSB_PARITY_REPORT
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
wc -l $TTSB_PARITY_REPORT
