"""Where a single request's milliseconds go (BASELINE configs[0]): the engine's per-category event timers over 20 warm steps.
    python tests/profile_latency.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint, workload  # noqa: E402

eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP16)
w = workload("single")
for _ in range(5):
    eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=3, padded_lstm=False)
plain = []
for _ in range(20):
    eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=3, padded_lstm=False)
    plain.append(eng.last_step_ms())
eng.profile(True)
tot = 0.0
for _ in range(20):
    eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=3, padded_lstm=False)
    tot += eng.last_step_ms()
p = eng.profile_read()
eng.profile(False)
plain.sort()
print(f"plain step {plain[len(plain) // 2]:.3f} ms (median of 20); with category timers {tot / 20:.3f} ms")
print("| category | ms per step | launches per step | us per launch |\n|---|---:|---:|---:|")
acc = 0.0
for k, v in p.items():
    if v["launches"]:
        acc += v["ms"] / 20
        print(f"| {k} | {v['ms'] / 20:.3f} | {v['launches'] / 20:.0f} | {1e3 * v['ms'] / v['launches']:.1f} |")
print(f"| sum | {acc:.3f} | | |")
