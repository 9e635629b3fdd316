#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "committed" 2>&1 | tail -3
