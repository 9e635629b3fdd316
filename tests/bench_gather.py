"""Rate of the duration-to-frame length-regulation gather (kmodel.py:116-139 as a row gather: no one-hot matrix, no GEMM) against
the measured HBM peak, at the step's size (32 utterances, ~78 tokens -> ~275 frames each, C = 512 and 640) and larger.
Algorithmic bytes: 4 C (T + F).  Best of 8.      python tests/bench_gather.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
hbm = float(json.load(open(pk)).get("hbm_gbs", 0)) if os.path.exists(pk) else 0.0
hbm = hbm or 6507.0
eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP16)
g = torch.Generator().manual_seed(1)
print(f"HBM peak used: {hbm:.0f} GB/s")
print("| utterances x tokens | C | frames | MB moved | best us | GB/s | % of HBM peak |\n|---|---:|---:|---:|---:|---:|---:|")
for B, T, C in ((32, 78, 512), (32, 78, 640), (128, 300, 640), (512, 300, 640), (1024, 500, 640)):
    dur = torch.randint(1, 7, (B * T,), generator=g, dtype=torch.int32).cuda()      # the op takes device pointers (include/ttsb200.h:161)
    off = [T * b for b in range(B + 1)]
    F = int(dur.sum())
    src = torch.randn(B * T, C, device="cuda")
    best = min(eng.op_length_regulate(dur, off, src, F, timed=True)[3] for _ in range(8))
    by = 4.0 * C * (B * T + F)
    print(f"| {B} x {T} | {C} | {F} | {by / 1e6:.0f} | {best * 1e3:.1f} | {by / best / 1e6:.0f} | {100 * by / best / 1e6 / hbm:.0f} |", flush=True)
    del src
