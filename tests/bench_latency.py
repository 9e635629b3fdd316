"""Single-request latency (BASELINE configs[0]: one 14-symbol utterance; and one 75-symbol utterance): device time of a step
(CUDA events, begin -> finish incl. the host gap), wall time of the call, kernels launched.  Median of 20 after 5 warm-ups.
    python tests/bench_latency.py            (env switches of the library apply, e.g. TTSB_UMMA_NOPACK=1)"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import random_ids, synthetic_checkpoint, synthetic_voice_pack, workload  # noqa: E402

eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP16)
g = torch.Generator().manual_seed(5)
cases = [("configs[0]: 14 symbols", workload("single")),
         ("75 symbols", dict(ids=[random_ids(g, 75)], ref_s=synthetic_voice_pack(1)[74], speeds=[1.0]))]
print("| request | audio-s | kernels | device ms | wall ms | audio-s/s (wall) |\n|---|---:|---:|---:|---:|---:|")
for name, w in cases:
    dev, wall, n = [], [], 0
    for i in range(25):
        l0 = eng.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=3, padded_lstm=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if i >= 5:
            dev.append(eng.last_step_ms()); wall.append((t1 - t0) * 1e3)
        n = eng.launch_count() - l0
    secs = sum(a.numel() for a in r.audio) / 24000.0
    d, wl = statistics.median(dev), statistics.median(wall)
    print(f"| {name} | {secs:.2f} | {n} | {d:.2f} | {wl:.2f} | {secs / wl * 1e3:.0f} |", flush=True)
