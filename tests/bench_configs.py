"""All five BASELINE.json configs on one B200 (bench.py measures configs[1] only; this fills profiles/rNN_configs.md).
    python tests/bench_configs.py [f16|bf16|fp32] > profiles/r01_configs.md
Numbers are audio-seconds per device-second (CUDA events on the engine's stream, median of the timed runs after warm-up)."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint, vocoder_inputs, workload  # noqa: E402

prec_name = sys.argv[1] if len(sys.argv) > 1 else "f16"
prec = {"fp32": pkg.PRECISION_FP32, "bf16": pkg.PRECISION_BF16, "f16": pkg.PRECISION_FP16}[prec_name]
eng = pkg.Engine(synthetic_checkpoint(1234), 0, prec)
print(f"# BASELINE configs on 1 x B200, precision {prec_name}, seeded synthetic weights\n")
print("| config | workload | audio-s per call | device ms (median) | audio-s/s | note |\n|---|---|---:|---:|---:|---|")


def run(ids, ref, speeds, reps=5, warm=2, padded=True):
    ms, secs = [], 0.0
    for i in range(warm + reps):
        r = eng.synthesize(ids, ref, speeds, seed=3, padded_lstm=padded)
        secs = sum(a.numel() for a in r.audio) / 24000.0
        if i >= warm:
            ms.append(eng.last_step_ms())
    return secs, statistics.median(ms)


# 1: single short utterance (latency)
w = workload("single")
secs, ms = run(w["ids"], w["ref_s"], w["speeds"], reps=10, padded=False)
print(f"| 1 | single 14-symbol utterance (B=1) | {secs:.2f} | {ms:.2f} | {secs / ms * 1e3:.0f} | latency case: {ms:.2f} ms per request |")
# 2: batch of 32
w = workload("batch32")
secs, ms = run(w["ids"], w["ref_s"], w["speeds"])
print(f"| 2 | 32 utterances x 60-90 symbols | {secs:.1f} | {ms:.2f} | {secs / ms * 1e3:.0f} | bench.py workload |")
# 3: long paragraph, 6 sequential chunks of 300-400 symbols, B=1 each (pipeline.py:400-420)
w = workload("paragraph")
tot_s, tot_ms = 0.0, 0.0
for i in range(6):
    s_, m_ = run([w["ids"][i]], w["ref_s"][i:i + 1], [1.0], reps=3, warm=1, padded=False)
    tot_s += s_
    tot_ms += m_
print(f"| 3 | ~500-char paragraph: 6 sequential chunks (B=1) | {tot_s:.1f} | {tot_ms:.2f} | {tot_s / tot_ms * 1e3:.0f} | chunks run one after another as the reference does |")
secs, ms = run(w["ids"], w["ref_s"], w["speeds"], reps=3, warm=1)
print(f"| 3b | same 6 chunks as ONE batch | {secs:.1f} | {ms:.2f} | {secs / ms * 1e3:.0f} | what a frame-budgeted batcher would do |")
# 4: vocoder sweep
for B in (1, 16, 64):
    for L in (100, 200, 400, 1000, 2000, 4000):           # the full grid
        if B * L > 64 * 1000:
            # keep the workspace bounded (64 x 4000 would need > 100 GB of fp32 stage buffers): the cell is listed, not run
            print(f"| 4 | vocoder only B={B} L={L} | {B * L * 300 / 24000.0:.1f} | - | - | not run: {B * L} generator frames exceed the 64 000-frame workspace bound of this sweep |")
            continue
        x, s, f0, noise = vocoder_inputs(B, L)
        ms_l = []
        for i in range(4):
            eng.vocoder(x, s, f0, None, seed=1)
            if i >= 1:
                ms_l.append(eng.last_step_ms())
        m_ = statistics.median(ms_l)
        secs = B * L * 300 / 24000.0
        print(f"| 4 | vocoder only B={B} L={L} | {secs:.1f} | {m_:.2f} | {secs / m_ * 1e3:.0f} | generator alone, in-kernel noise |")
        del x, noise
# 5: 512 utterances of the 4096-utterance fleet, micro-batched by a frame budget
w = workload("fleet512")
cost = [len(x) for x in w["ids"]]
mbs = pkg.micro_batches(cost, budget=3500)
t_ms, t_s = 0.0, 0.0
for it in range(2):
    t_ms, t_s = 0.0, 0.0
    for mb in mbs:
        r = eng.synthesize([w["ids"][i] for i in mb], w["ref_s"][mb], [w["speeds"][i] for i in mb], seed=5)
        t_ms += eng.last_step_ms()
        t_s += sum(a.numel() for a in r.audio) / 24000.0
print(f"| 5 | 512 of the 4096 fleet utterances, 16 voices, speeds 0.8-1.2, {len(mbs)} micro-batches (token budget 3500) | {t_s:.0f} | {t_ms:.1f} | {t_s / t_ms * 1e3:.0f} | one GPU's share; ranks are independent (bench.py --gpus N) |")
