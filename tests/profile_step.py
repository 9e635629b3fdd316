"""One warm-up step + one step of BASELINE configs[1] (batch of 32) — the command line ncu wraps (profiles/README.md).
    python tests/profile_step.py [bf16|fp32] [n_utt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint, workload  # noqa: E402

prec = {"fp32": pkg.PRECISION_FP32, "bf16": pkg.PRECISION_BF16, "f16": pkg.PRECISION_FP16}[sys.argv[1] if len(sys.argv) > 1 else "f16"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = pkg.Engine(synthetic_checkpoint(1234), 0, prec)
w = workload("batch32")
for step in range(2):
    l0 = eng.launch_count()
    r = eng.synthesize(w["ids"][:n], w["ref_s"][:n], w["speeds"][:n], seed=7)
    torch.cuda.synchronize()
    print(f"step {step}: {eng.last_step_ms():.2f} ms, {eng.launch_count() - l0} launches, {sum(a.numel() for a in r.audio) / 24000:.1f} audio-s", flush=True)
