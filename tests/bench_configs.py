"""All five BASELINE.json configs on one B200, next to the two baselines the survey names (bench.py measures configs[1] only):
    python tests/bench_configs.py [f16|bf16|fp32] [--no-baselines] > profiles/rNN_configs.md
Engine numbers are audio-seconds per device-second (CUDA events on the engine's stream, median of the timed runs after warm-up).
Baselines, on the SAME inputs (bounded samples where stated): the CPU oracle (fp32 PyTorch restatement of the reference graph, all
host cores) and the same module moved to the GPU and run in PyTorch eager (stock cuDNN / cuBLAS kernels, per-sample decode loop)."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint, vocoder_inputs, workload  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
prec_name = args[0] if args else "f16"
baselines = "--no-baselines" not in sys.argv
prec = {"fp32": pkg.PRECISION_FP32, "bf16": pkg.PRECISION_BF16, "f16": pkg.PRECISION_FP16}[prec_name]
eng = pkg.Engine(synthetic_checkpoint(1234), 0, prec)
cpu = gpu = None
if baselines:
    from oracle.synth_weights import build_oracle
    cpu = build_oracle(1234)
    gpu = build_oracle(1234).cuda()
    torch.set_num_threads(os.cpu_count() or 1)
cores = os.cpu_count() or 1
print(f"# BASELINE configs on 1 x B200, precision {prec_name}, seeded synthetic weights\n")
print(f"CPU oracle on {cores} host cores (fp32 PyTorch eager); GPU eager = the same oracle module on cuda:0 (stock PyTorch kernels). "
      "Baselines run once after a warm-up call, on the sample named in the row.\n")
print("| config | workload | audio-s per call | engine device ms (median) | engine audio-s/s | CPU oracle audio-s/s | B200 PyTorch eager audio-s/s | engine / eager | note |")
print("|---|---|---:|---:|---:|---:|---:|---:|---|")


def run(ids, ref, speeds, reps=5, warm=2, padded=True):
    ms, secs = [], 0.0
    for i in range(warm + reps):
        r = eng.synthesize(ids, ref, speeds, seed=3, padded_lstm=padded)
        secs = sum(a.numel() for a in r.audio) / 24000.0
        if i >= warm:
            ms.append(eng.last_step_ms())
    return secs, statistics.median(ms)


def oracle_rate(m, fn):
    """audio-s/s of one call of fn(m) after a small warm-up call made by the caller"""
    cuda = next(m.parameters()).is_cuda
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        out = fn(m)
    if cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return sum(a.numel() for a in out) / 24000.0 / dt


def base_cols(fn, engine_rate):
    if not baselines:
        return "- | - | -"
    for m in (cpu, gpu):                              # lazy initialisation (cuDNN plans, thread pools) out of the timing
        with torch.no_grad():
            m.forward_batch(w1["ids"], w1["ref_s"], w1["speeds"], noise_seed=0)
    c, g = oracle_rate(cpu, fn), oracle_rate(gpu, fn)
    return f"{c:.1f} | {g:.1f} | {engine_rate / g:.0f}x"


def fwd(w, n=None):
    return lambda m: m.forward_batch(w["ids"][:n], w["ref_s"][:n], w["speeds"][:n], noise_seed=0).audio


# 1: single short utterance (latency)
w1 = workload("single")
secs, ms = run(w1["ids"], w1["ref_s"], w1["speeds"], reps=10, padded=False)
print(f"| 1 | single 14-symbol utterance (B=1) | {secs:.2f} | {ms:.2f} | {secs / ms * 1e3:.0f} | {base_cols(fwd(w1), secs / ms * 1e3)} | latency case: {ms:.2f} ms per request |")
# 2: batch of 32
w = workload("batch32")
secs, ms = run(w["ids"], w["ref_s"], w["speeds"])
print(f"| 2 | 32 utterances x 60-90 symbols | {secs:.1f} | {ms:.2f} | {secs / ms * 1e3:.0f} | {base_cols(fwd(w), secs / ms * 1e3)} | bench.py workload; baselines on all 32 |")
# 3: long paragraph, 6 sequential chunks of 300-400 symbols, B=1 each (pipeline.py:400-420)
w = workload("paragraph")
tot_s, tot_ms = 0.0, 0.0
for i in range(6):
    s_, m_ = run([w["ids"][i]], w["ref_s"][i:i + 1], [1.0], reps=3, warm=1, padded=False)
    tot_s += s_
    tot_ms += m_


def seq_chunks(m):
    out = []
    for i in range(6):
        out += m.forward_batch([w["ids"][i]], w["ref_s"][i:i + 1], [1.0], noise_seed=0).audio
    return out


print(f"| 3 | ~500-char paragraph: 6 sequential chunks (B=1) | {tot_s:.1f} | {tot_ms:.2f} | {tot_s / tot_ms * 1e3:.0f} | {base_cols(seq_chunks, tot_s / tot_ms * 1e3)} | chunks run one after another as the reference does |")
secs, ms = run(w["ids"], w["ref_s"], w["speeds"], reps=3, warm=1)
print(f"| 3b | same 6 chunks as ONE batch | {secs:.1f} | {ms:.2f} | {secs / ms * 1e3:.0f} | - | - | - | what a frame-budgeted batcher would do |")
# 4: vocoder sweep (the full grid of SURVEY 8(d) config 4)
BASE_CELLS = {(1, 100), (1, 1000), (1, 4000), (16, 1000)}
for B in (1, 16, 64):
    for L in (100, 200, 400, 1000, 2000, 4000):
        x, s, f0, noise = vocoder_inputs(B, L)
        ms_l = []
        try:
            for i in range(3):
                eng.vocoder(x, s, f0, None, seed=1)
                if i >= 1:
                    ms_l.append(eng.last_step_ms())
        except Exception as exc:
            print(f"| 4 | vocoder only B={B} L={L} | {B * L * 300 / 24000.0:.1f} | - | - | - | - | - | failed: {str(exc)[:80]} |")
            continue
        m_ = statistics.median(ms_l)
        secs = B * L * 300 / 24000.0
        cols = "- | - | -"
        if baselines and (B, L) in BASE_CELLS:
            def voc(m, x=x, s=s, f0=f0, noise=noise):
                d = next(m.parameters()).device
                return list(m.vocoder(x.to(d), s.to(d), f0.to(d), noise.to(d)))
            cols = base_cols(voc, secs / m_ * 1e3)
        print(f"| 4 | vocoder only B={B} L={L} | {secs:.1f} | {m_:.2f} | {secs / m_ * 1e3:.0f} | {cols} | generator alone, in-kernel noise |")
        del x, noise
# 5: 512 utterances of the 4096-utterance fleet, micro-batched by a token budget
w = workload("fleet512")
cost = [len(x) for x in w["ids"]]
mbs = pkg.micro_batches(cost, budget=3500)
t_ms, t_s = 0.0, 0.0
for it in range(2):
    t_ms, t_s = 0.0, 0.0
    for mb in mbs:
        r = eng.synthesize([w["ids"][i] for i in mb], w["ref_s"][mb], [w["speeds"][i] for i in mb], seed=5, padded_lstm=False)
        t_ms += eng.last_step_ms()
        t_s += sum(a.numel() for a in r.audio) / 24000.0
print(f"| 5 | 512 of the 4096 fleet utterances, 16 voices, speeds 0.8-1.2, {len(mbs)} micro-batches (token budget 3500) | {t_s:.0f} | {t_ms:.1f} | {t_s / t_ms * 1e3:.0f} | "
      f"{base_cols(fwd(w, 32), t_s / t_ms * 1e3)} | one GPU's share (bench.py --workload fleet --gpus N runs the whole fleet); baselines on the first 32 utterances |")
