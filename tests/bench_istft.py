"""Streaming rate of the fused exp/sin + iSTFT(20,5) kernel against the measured HBM peak, at the in-step size (32 utterances
x ~275 decoder frames) and at larger sizes: separates launch ramp / tail of a 40-us kernel from its steady-state rate.
Algorithmic bytes: 22 channels read per STFT frame + 600 samples written per decoder frame, fp32.  Best of 8.
    python tests/bench_istft.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", peaks.get("hbm_GBps", 0)) or 0) or 6507.0
eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP16)
g = torch.Generator(device="cuda").manual_seed(1)
print(f"HBM peak used: {hbm:.0f} GB/s")
print("| utterances x decoder frames | MB moved | best us | GB/s | % of HBM peak |\n|---|---:|---:|---:|---:|")
for B, F in ((32, 275), (32, 1100), (128, 1100), (128, 4400)):
    Fs = [F] * B
    rows = sum(120 * f + 1 for f in Fs)
    cp = torch.randn(rows, 24, device="cuda", generator=g) * 0.5
    by = 4.0 * (22 * rows + 600 * sum(Fs))
    best = min(eng.op_istft(cp, Fs, timed=True)[1] for _ in range(8))
    print(f"| {B} x {F} | {by / 1e6:.0f} | {best * 1e3:.1f} | {by / best / 1e6:.0f} | {100 * by / best / 1e6 / hbm:.0f} |", flush=True)
    del cp
