#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_lengths or two_conv" 2>&1 | tail -5
python - <<'PY'
import os, subprocess, sys, tempfile, torch
sys.path.insert(0, os.getcwd())
from tests.test_gpu_parity import _AB_SCRIPT
from tests.parity import snr_db
d = tempfile.mkdtemp()
open(d + "/ab.py", "w").write(_AB_SCRIPT.format(root=os.getcwd()))
outs = {}
for tag, extra in (("wa", {}), ("wa1", {"TTSB_NO_WA0": "1"}), ("m", {"TTSB_NO_WA": "1"})):
    env = dict(os.environ); env.update(extra)
    subprocess.run([sys.executable, d + "/ab.py", f"{d}/{tag}.pt"], env=env, check=True)
    outs[tag] = torch.load(f"{d}/{tag}.pt")
for t in ("wa", "wa1"):
    print(t, "vs time-as-M kernel:", [round(snr_db(m, a), 1) for a, m in zip(outs[t], outs["m"])])
PY
