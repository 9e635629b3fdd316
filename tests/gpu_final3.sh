#!/bin/bash
# last check of the committed tree: the driver's own sequence (GPU tests, smoke, both bench arms with default flags)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference 2> gpurun_out/ref3.err | tail -1 | cut -c1-300
timeout 900 python bench.py 2> gpurun_out/bench3.err | tee gpurun_out/bench3.json | tail -1 | cut -c1-400
