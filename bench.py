#!/usr/bin/env python
"""bench.py — audio-seconds synthesised per second (RTF^-1) of the synthesis hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f16|bf16|fp32] [--impl b200|reference] [--workload batch32|fleet]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path (BatchKModel.forward_batch semantics: phoneme ids + voice style rows -> 24 kHz
waveform) over one batch of synthetic utterances.  Workload at every N: BASELINE.json configs[1] per GPU — 32 utterances of
60..90 symbols (~20 Chinese characters), one voice; rank r uses seed 2+r, so per-GPU work is fixed ("weak" scaling) and
utterances are partitioned by rank with no collective on the data path.

--workload fleet is BASELINE.json configs[4]: 4096 utterances, 16 voice packs, speeds 0.8-1.2, dealt to the N ranks by LPT on the
predicted length and run as token-budgeted micro-batches - total work fixed ("strong" scaling); a step is one pass over the fleet.

value  = audio-seconds / device time, CUDA events on the engine's stream spanning each step (max over ranks);
e2e    = the same metric through the reference-facing class (BatchKModel.forward_batch + .cpu() of every waveform):
         host strings / pinned style rows in, host PCM out, wall clock with a synchronize on both sides.
--impl reference times the CPU oracle (the reference's PyTorch CPU path restated, oracle/kokoro_oracle.py) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio-seconds synthesized per second (RTF^-1)"
UNIT = "audio-s/s"
SR = 24000.0
FLOP_PER_FRAME = 1.3169e9          # 658.45 MMAC per 25 ms decoder frame
FLOP_PER_TOKEN_BASE = 2 * 79.42e6  # + 2*0.018432e6*T per token (ALBERT attention)


def algorithmic_flops(tokens, frames):
    return sum(t * (FLOP_PER_TOKEN_BASE + 2 * 0.018432e6 * t) for t in tokens) + FLOP_PER_FRAME * sum(frames)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of THIS rank's GPU, sampled every 25 ms while the timed region runs.
    In-process NVML: initialised before the region, a sample costs microseconds, so even a quarter-second region (10 steps)
    gets a dozen samples.  (A spawned `nvidia-smi -lms 200` delivered its first row only after its own start-up - on an
    8-GPU box later than the end of a short region: samples = 0.)  Falls back to nvidia-smi when NVML cannot be used."""
    REASONS = (("hw_slowdown", ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"), 0x8),
               ("hw_thermal_slowdown", ("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"), 0x40),
               ("sw_thermal_slowdown", ("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), 0x20),
               ("sw_power_cap", ("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"), 0x4))
    PERIOD = 0.025

    def __init__(self, index, nvml=None, uuid=None):
        self.index, self.uuid, self.nvml = index, uuid, nvml
        self.sm, self.mx, self.reasons = [], [], set()
        self.h = self.t = self.smi = None
        self._stop = threading.Event()
        self.source = None

    # ---- NVML ----
    def _open_nvml(self):
        nv = self.nvml
        if nv is None:
            import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            for cand in (self.uuid, self.uuid.encode()):
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(cand)
                    break
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.nvml, self.h = nv, h
        self.masks = [(name, next((getattr(nv, a) for a in attrs if hasattr(nv, a)), default)) for name, attrs, default in self.REASONS]
        self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self._sample()                                       # must work once, or we fall back

    def _sample(self):
        nv = self.nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        bits = int(self.get_reasons(self.h))
        for name, mask in self.masks:
            if bits & mask:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.wait(self.PERIOD):
            try:
                self._sample()
            except Exception:
                return

    # ---- nvidia-smi fallback ----
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def _read_smi(self):
        for line in self.smi.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def start(self):
        """call right before the timed region (NVML is opened here if `open()` was not called earlier)"""
        if self.source is None:
            self.open()
        if self.source == "nvml":
            self.sm.clear(); self.mx.clear(); self.reasons.clear()     # the probe sample of open() was taken outside the region
            self._sample()
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def open(self):
        """the slow part (NVML initialisation or the start-up of nvidia-smi), outside the timed region"""
        try:
            self._open_nvml()
            self.source = "nvml"
        except Exception:
            self.h = None
            try:
                self.smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read_smi, daemon=True)
                self.t.start()
                self.source = "nvidia-smi"
            except Exception:
                self.source = "none"

    def stop(self):
        if self.source == "nvml":
            try:
                self._sample()                               # the region's last moment
            except Exception:
                pass
            self._stop.set()
            if self.t:
                self.t.join(timeout=1)
        elif self.source == "nvidia-smi":
            self.smi.terminate()
            try:
                self.smi.wait(timeout=2)
            except Exception:
                self.smi.kill()
        else:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no NVML and no nvidia-smi"], samples=0, source="none")
        return dict(sm_mhz=statistics.median(self.sm) if self.sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                    reasons=sorted(self.reasons), samples=len(self.sm), source=self.source)


def make_workload(rank: int):
    from illufly_tts_b200.synthetic import ids_to_phonemes, workload     # seeded synthetic inputs (plain data, no model code)
    w = workload("batch32", seed=2 + rank)
    w["phonemes"] = [ids_to_phonemes(ids) for ids in w["ids"]]
    return w


def make_fleet(n: int, rank: int, world: int, token_budget: int):
    """BASELINE configs[4]: n utterances (4096), 16 voice packs, speeds U(0.8, 1.2), the SAME fleet on every rank; dealt to the ranks
    by LPT on the predicted length (tokens / speed), this rank's share cut into token-budgeted micro-batches of similar lengths."""
    import illufly_tts_b200 as pkg
    from illufly_tts_b200.synthetic import ids_to_phonemes, workload
    w = workload(f"fleet{n}")
    cost = [len(x) / s for x, s in zip(w["ids"], w["speeds"])]
    mine = pkg.lpt_partition(cost, world)[rank]
    tok = [len(w["ids"][i]) for i in mine]
    mbs = [[mine[j] for j in mb] for mb in pkg.micro_batches(tok, budget=token_budget)]
    w["phonemes"] = {i: ids_to_phonemes(w["ids"][i]) for i in mine}
    return w, mine, mbs, sum(cost[i] for i in mine), sum(cost)


WORKLOADS = {"batch32": "configs[1]: batch of 32 short utterances (60-90 symbols each), one voice, per GPU",
             "fleet": "configs[4]: fleet of {n} utterances (40-120 symbols), 16 voice packs, speeds 0.8-1.2, partitioned over the GPUs by LPT, token-budgeted micro-batches"}


def config_for(args, tokens, frames, utterances):
    """the SAME dict in both arms (the driver compares them): facts of the workload only, nothing about how an arm ran it"""
    secs = None if frames is None else frames * 600 / SR
    if args.workload == "fleet":
        return dict(workload=WORKLOADS["fleet"].format(n=args.fleet_size), utterances_total=utterances, tokens_total=tokens, frames_total=frames,
                    audio_s_per_step=secs, weights="seeded random init, Kokoro-82M architecture", token_budget=args.token_budget,
                    l2="inputs larger than L2: every stage buffer of a micro-batch exceeds 126 MB")
    return dict(workload=WORKLOADS["batch32"], utterances_per_gpu=utterances, tokens_per_gpu=tokens, frames_per_gpu=frames,
                audio_s_per_step_per_gpu=secs, weights="seeded random init, Kokoro-82M architecture",
                partitioning="utterances by rank, no collective", l2="inputs larger than L2: every stage buffer of a step exceeds 126 MB")


def cpu_oracle_rate(m, w, n_utt, threads):
    import torch
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    o = m.forward_batch(w["ids"][:n_utt], w["ref_s"][:n_utt], w["speeds"][:n_utt], noise_seed=0)
    dt = time.perf_counter() - t0
    secs = sum(a.numel() for a in o.audio) / SR
    return secs / dt, secs, dt


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (its PyTorch graph, restated in oracle/) on the host cores: the whole batch of
    configs[1] per step (fleet: the first 32 utterances of the fleet, a bounded sample), all host threads."""
    if rank != 0:
        return
    import torch
    from oracle.synth_weights import build_oracle
    threads = os.cpu_count() or 1
    m = build_oracle(1234)
    if args.workload == "fleet":
        from illufly_tts_b200.synthetic import workload
        w = workload(f"fleet{args.fleet_size}")
        tokens_all = sum(len(x) for x in w["ids"])
    else:
        w = make_workload(0)
    n_utt = min(args.ref_utts, len(w["ids"]))
    n_warm = min(max(args.warmup, 0), 1)       # a CPU step is seconds of work: one warm-up step (lazy initialisation), at most 3 timed
    for _ in range(n_warm):
        cpu_oracle_rate(m, w, 1, threads)
    rates, tot_s, tot_t = [], 0.0, 0.0
    for _ in range(max(1, min(args.steps, 3))):
        r, secs, dt = cpu_oracle_rate(m, w, n_utt, threads)
        rates.append(r); tot_s += secs; tot_t += dt
    value = tot_s / tot_t
    frames = int(round(tot_s / len(rates) * SR / 600))
    if args.workload == "fleet":
        # frames of the whole fleet are not known without running it: extrapolate from the sample (reported as such)
        cfg = config_for(args, tokens_all, None, len(w["ids"]))
        sample = f"first {n_utt} of the {len(w['ids'])} fleet utterances per step ({tot_s / len(rates):.1f} audio-s)"
    else:
        cfg = config_for(args, sum(len(x) for x in w["ids"]), frames if n_utt >= len(w["ids"]) else None, len(w["ids"]))
        sample = f"{'all' if n_utt >= len(w['ids']) else 'first ' + str(n_utt) + ' of the'} {len(w['ids'])} utterances of configs[1] per step ({tot_s / len(rates):.1f} audio-s)"
    sample += ", fp32 PyTorch CPU eager, per-sample decode loop as kmodel.py:108-144"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(rates), warmup=n_warm, ms_per_step=1e3 * tot_t / len(rates),
                higher_is_better=True, scaling="strong" if args.workload == "fleet" else "weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=cfg, requested=dict(steps=args.steps, warmup=args.warmup), bounded="steps <= 3, warmup <= 1: every CPU step is seconds of work",
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def gpu_eager_rate(w, n_utt):
    """The real competitor: the SAME graph through stock PyTorch on this GPU - the oracle module moved to CUDA, eager
    mode, cuDNN / cuBLAS kernels, PyTorch's default math modes (TF32 allowed in cuDNN convolutions, not in matmuls), the reference's
    own per-sample decode loop (kmodel.py:108-144).  Outside every timed region of this bench; one warm-up pass, one timed pass."""
    import torch
    from oracle.synth_weights import build_oracle
    m = build_oracle(1234).cuda()
    with torch.no_grad():
        m.forward_batch(w["ids"][:2], w["ref_s"][:2], w["speeds"][:2], noise_seed=0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o = m.forward_batch(w["ids"][:n_utt], w["ref_s"][:n_utt], w["speeds"][:n_utt], noise_seed=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    secs = sum(a.numel() for a in o.audio) / SR
    del m, o
    torch.cuda.empty_cache()
    return dict(value=secs / dt, unit=UNIT, utterances=n_utt, audio_s=secs, seconds=dt,
                what="oracle/kokoro_oracle.py on cuda:0, PyTorch eager fp32 (cudnn.allow_tf32 = %s, matmul.allow_tf32 = %s), per-sample decode loop; "
                     "measured once after a warm-up pass, outside the timed region" % (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32))


def hbm_standalone(eng, torch):
    """standalone HBM-bound kernels on buffers larger than L2 (north star: fused elementwise / iSTFT against the HBM peak)"""
    hbm_ops = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, C = 1 << 20, 128                                     # 512 MB in, 512 MB out
    xb = torch.randn(rows, C, device="cuda", generator=g)
    gb = torch.randn(4, 2 * C, device="cuda", generator=g) * 0.2
    al = torch.rand(C, device="cuda", generator=g) + 0.5
    offs = [0, rows // 4, rows // 2, 3 * rows // 4, rows]
    best = (1e9, 1e9)
    for _ in range(4):
        _, ms_s, ms_a = eng.op_adain_snake(xb, offs, gb, al, timed=True)
        best = (min(best[0], ms_s), min(best[1], ms_a))
    hbm_ops["adain_snake_apply"] = dict(bytes=8.0 * rows * C, ms=best[1])
    hbm_ops["instance_stats"] = dict(bytes=4.0 * rows * C, ms=best[0])
    Fs = [2200] * 4                                             # 4 x 264001 STFT frames
    cp = torch.randn(sum(120 * f + 1 for f in Fs), 24, device="cuda", generator=g) * 0.5
    bi = 1e9
    for _ in range(4):
        _, ms_i = eng.op_istft(cp, Fs, timed=True)
        bi = min(bi, ms_i)
    hbm_ops["istft_standalone"] = dict(bytes=4.0 * (22 * cp.shape[0] + 600 * sum(Fs)), ms=bi)      # the step's amount of data: ~half ramp and tail
    del xb, cp
    Fs = [1100] * 128                                           # 16x the step's data (1.8 GB): the kernel's steady-state rate
    cp = torch.randn(sum(120 * f + 1 for f in Fs), 24, device="cuda", generator=g) * 0.5
    bi = 1e9
    for _ in range(4):
        _, ms_i = eng.op_istft(cp, Fs, timed=True)
        bi = min(bi, ms_i)
    hbm_ops["istft_streaming_1p8GB"] = dict(bytes=4.0 * (22 * cp.shape[0] + 600 * sum(Fs)), ms=bi)
    del cp
    # length-regulation gather at a streaming size (512 utterances x 300 tokens, C = 640: 1.8 GB)
    Bg, Tg, Cg = 512, 300, 640
    dur = torch.randint(1, 7, (Bg * Tg,), generator=torch.Generator().manual_seed(3), dtype=torch.int32).cuda()
    srcg = torch.randn(Bg * Tg, Cg, device="cuda", generator=g)
    Fg = int(dur.sum())
    bg = 1e9
    for _ in range(4):
        bg = min(bg, eng.op_length_regulate(dur, [Tg * b for b in range(Bg + 1)], srcg, Fg, timed=True)[3])
    hbm_ops["gather_streaming_1p8GB"] = dict(bytes=4.0 * Cg * (Bg * Tg + Fg), ms=bg)
    del srcg, dur
    return hbm_ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 200 (batch32: ~5 s of timed region, sustained clocks) / 3 (fleet: a step is the whole fleet)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TTSB_BENCH_PRECISION", "f16"), choices=["f16", "bf16", "fp32"])
    ap.add_argument("--ref-utts", type=int, default=32, help="utterances per step of the CPU reference arm (32 = the whole batch of configs[1])")
    ap.add_argument("--cpu-baseline-utts", type=int, default=32)
    ap.add_argument("--workload", default="batch32", choices=["batch32", "fleet"])
    ap.add_argument("--fleet-size", type=int, default=4096)
    ap.add_argument("--token-budget", type=int, default=3500, help="fleet: tokens per micro-batch (bounds the engine's workspace)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the stock-PyTorch-on-GPU baseline (rank 0, N=1, outside the timed region)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-ops", action="store_true", help="skip the standalone HBM-kernel measurements")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if (args.workload == "fleet" or args.impl == "reference") else 200
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import illufly_tts_b200 as pkg
    from illufly_tts_b200.synthetic import synthetic_checkpoint, synthetic_config
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3) if args.workload == "batch32" else max(args.warmup, 1)
    prec = {"f16": pkg.PRECISION_FP16, "bf16": pkg.PRECISION_BF16, "fp32": pkg.PRECISION_FP32}[args.precision]
    tier = {"f16": "tcgen05, fp16 operands, fp32 accumulate (the north star's >= 40 dB tier: 59.1 dB measured on bit-identical generator inputs)",
            "bf16": "tcgen05, bf16 operands, fp32 accumulate (>= 40 dB tier: 42.8 dB measured)",
            "fp32": "the strict tier (max-abs 1e-3): tcgen05, split-fp16 operands (x = hi + lo, 3 MMAs per K step), fp32 accumulate"}[args.precision]

    ckpt = synthetic_checkpoint(1234)                                   # random-init weights of the Kokoro-82M architecture
    fleet = args.workload == "fleet"
    # the fleet runs with per-utterance semantics (padded_lstm=False): an utterance's durations must not depend on its micro-batch
    model = pkg.BatchKModel(config=synthetic_config(), model=ckpt, precision=prec, strict=True, padded_lstm=not fleet).to(f"cuda:{local}").eval()
    eng = model.engine
    if fleet:
        w, mine, mbs, my_cost, all_cost = make_fleet(args.fleet_size, rank, world, args.token_budget)
        batches = [dict(ids=[w["ids"][i] for i in mb], ref_s=w["ref_s"][mb], speeds=[w["speeds"][i] for i in mb],
                        phonemes=[w["phonemes"][i] for i in mb]) for mb in mbs]
        n_fleet_utts, tokens_all = len(w["ids"]), sum(len(x) for x in w["ids"])
    else:
        w = make_workload(rank)
        batches = [w]
    for b in batches:
        b["ref_pinned"] = b["ref_s"].unsqueeze(1).contiguous().pin_memory()      # [B,1,256] like pipeline.py:603
    tokens = [len(x) for b in batches for x in b["ids"]]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(collect=None):
        """one pass of this rank's work; returns device ms (CUDA events on the engine's stream, summed over its micro-batches)"""
        ms = 0.0
        frames = []
        for b in batches:
            r = eng.synthesize(b["ids"], b["ref_s"], b["speeds"], noise=None, seed=7, padded_lstm=not fleet)
            ms += eng.last_step_ms()
            if collect is not None:
                for k, v in eng.profile_read().items():
                    a = collect.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                    for kk in a:
                        a[kk] += v[kk]
            frames += [int(a.numel()) // 600 for a in r.audio]
        return ms, frames

    # ---- device-timed steps (value) ----
    eng.profile(True)
    uid = getattr(torch.cuda.get_device_properties(local), "uuid", None)
    clocks = ClockSampler(local, uuid=("GPU-" + str(uid)) if uid is not None else None)
    clocks.open()                                          # NVML initialisation: before the warm-up, far from the timed region
    frames = None
    for _ in range(warmup):
        _, frames = one_step()
    audio_s = sum(frames) * 600 / SR
    sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms, prof_acc = 0.0, {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, _ = one_step(prof_acc)
        dev_ms += ms
    sync_all()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    eng.profile(False)

    hbm_ops = hbm_standalone(eng, torch) if (rank == 0 and not args.no_hbm_ops and not fleet) else {}

    # ---- end to end through the reference-facing class (e2e): host strings + pinned style rows in, host PCM out ----
    def e2e_step(pinned_out):
        n = 0
        for b in batches:
            if pinned_out:      # the engine's own pinned staging buffer (one D2H of the whole batch): Engine.synthesize(audio_to_host=True)
                ids = [pkg.kmodel.phonemes_to_ids(model.vocab, ph) for ph in b["phonemes"]]
                r = eng.synthesize(ids, b["ref_pinned"], b["speeds"], audio_to_host=True, padded_lstm=not fleet)
                n += sum(a.numel() for a in r.audio)
            else:               # what the reference's callers do (service.py:316, 380): forward_batch, then .cpu() per waveform
                outs = [a.cpu() for a in model.forward_batch(b["phonemes"], b["ref_pinned"], b["speeds"])]
                n += sum(o.numel() for o in outs)
        return n
    e2e_steps = args.steps if not fleet else 1
    e2e_res = {}
    for pinned_out in (False, True):
        for _ in range(2 if not fleet else 0):
            e2e_step(pinned_out)
        sync_all()
        t1 = time.perf_counter()
        n_samples = 0
        for _ in range(e2e_steps):
            n_samples = e2e_step(pinned_out)
        sync_all()
        e2e_res[pinned_out] = time.perf_counter() - t1
    d2h = n_samples * 4 + sum(tokens) * 8
    h2d = sum(tokens) * 4 + sum(len(b["ids"]) + 1 for b in batches) * 4 + sum(b["ref_pinned"].numel() for b in batches) * 4 + len(tokens) * 4

    # ---- max over ranks, totals over ranks ----
    t = torch.tensor([dev_ms, wall, e2e_res[False], e2e_res[True]], dtype=torch.float64, device="cuda")
    tot = torch.tensor([audio_s, float(launches), float(sum(frames)), float(sum(tokens))], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device="cuda")
    per_rank[rank] = torch.tensor([dev_ms / args.steps, audio_s, float(len(batches))], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max, e2e_max, e2e_pin_max = t.tolist()
    audio_total, launches_total, frames_total, tokens_total = tot.tolist()

    if rank == 0:
        peaks = load_peaks()
        dom = "conv_tensor" if prof_acc.get("conv_tensor", {}).get("launches", 0) else "conv_fp32"
        p = prof_acc[dom]
        ach = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0.0
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if dom == "conv_tensor" and args.precision == "f16" and os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic, traffic_note = tj["dram_bytes_per_launch_mean"], tj["source"]
        sm_clk = clk.get("sm_mhz") or 0.0
        roofline = dict(bound="tensor", kernel="conv_umma_kernel + conv_umma_wa_kernel (tcgen05 implicit-GEMM Conv1d, every convolution / projection of the step)"
                        if dom == "conv_tensor" else "conv_simt_kernel (fp32 FFMA implicit-GEMM Conv1d)",
                        achieved=ach, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=ach / peaks["tf_sustained"],
                        # the timed region is seconds long at a power-capped clock: the sustained cuBLAS figure is the matching peak; the
                        # burst figure (a kernel timed alone at full clock) is given beside it
                        frac_of_burst_peak=ach / peaks["tf_burst"], peak_burst=peaks["tf_burst"],
                        traffic=traffic, traffic_note=traffic_note,
                        peak_source=f"bf16 sustained, {peaks['source']}", launches_per_step=p["launches"] / args.steps,
                        # the same kernels on the other axis: algorithmic bytes (inputs + outputs + residuals, from the launch arguments) per second
                        hbm_achieved_gbs=(p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0),
                        hbm_frac=(p["bytes"] / (p["ms"] * 1e-3) / 1e9 / peaks["hbm"] if p["ms"] > 0 else 0.0),
                        avg_launch_ms=p["ms"] / max(p["launches"], 1), share_of_step=p["ms"] / max(dev_ms, 1e-9),
                        algorithmic="FLOPs = 2 rows Cin Cout k per launch, summed from the launch arguments (3 MMAs per K step of the split-fp16 layers count once)")
        hbm = {}
        for k in ("istft", "upsample_add", "gather"):
            q = prof_acc.get(k)
            if q and q["ms"] > 0:
                g = q["bytes"] / (q["ms"] * 1e-3) / 1e9
                hbm[k] = dict(achieved_gbs=g, frac=g / peaks["hbm"], ms_per_step=q["ms"] / args.steps)
        for k, v in hbm_ops.items():
            gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            hbm[k] = dict(achieved_gbs=gbs, frac=gbs / peaks["hbm"], ms=v["ms"], note="standalone op, buffers > L2, best of 4")
        if fleet:
            cfg = config_for(args, int(tokens_all), int(frames_total), n_fleet_utts)
        else:
            cfg = config_for(args, sum(tokens), sum(frames), len(tokens))
        line = dict(metric=METRIC, value=audio_total * args.steps / (dev_ms_max * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=warmup,
                    ms_per_step=dev_ms_max / args.steps, higher_is_better=True, scaling="strong" if fleet else "weak", vs_baseline=None,
                    dtype={"f16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], data="synthetic", config=cfg,
                    precision=dict(tier=tier, text_side="split-fp16 operands on tcgen05 (fp32-class; decides the integer durations)", frame_side=args.precision),
                    wall_ms_per_step=1e3 * wall_max / args.steps,
                    e2e=dict(value=audio_total * e2e_steps / e2e_max, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps,
                             api="BatchKModel.forward_batch(list[str], pinned ref_s [B,1,256], speeds) + .cpu() per waveform (pageable copies, as service.py:316 does)",
                             # the same region with the engine's pinned staging buffer: ONE D2H copy of the batch.  Not the default of the drop-in class
                             # because its result would alias an engine-owned buffer, and the reference's callers keep what they get (pipeline.py:829)
                             value_pinned_host_buffer=audio_total * e2e_steps / e2e_pin_max),
                    workspace=dict(peak_bytes=eng.workspace_peak_bytes(), per_frame_bytes=eng.workspace_peak_bytes() / max(sum(frames[-len(batches[-1]["ids"]):]), 1),
                                   note="high-water mark of the engine's stack-allocated workspace in the last step (last micro-batch for the fleet)"),
                    gpu_launches=int(launches_total), clocks=clk, roofline=roofline, hbm_kernels=hbm,
                    profile_ms_per_step={k: v["ms"] / args.steps for k, v in prof_acc.items()},
                    # share of the device-timed step covered by the event-timed categories (the rest: small kernels, host gap between the halves)
                    profile_accounted_frac=sum(v["ms"] for v in prof_acc.values()) / max(dev_ms, 1e-9))
        line["algorithmic_tflops_per_s"] = (FLOP_PER_FRAME * frames_total + FLOP_PER_TOKEN_BASE * tokens_total) / (dev_ms_max / args.steps * 1e-3) / 1e12
        if fleet:
            pr = per_rank.tolist()
            busy = [x[0] for x in pr]
            line["fleet"] = dict(per_rank_device_ms=busy, per_rank_audio_s=[x[1] for x in pr], per_rank_micro_batches=[int(x[2]) for x in pr],
                                 imbalance_max_over_mean=max(busy) / (sum(busy) / len(busy)),
                                 partition="LPT on tokens / speed (the predicted length), then micro-batches of similar lengths within a token budget",
                                 semantics="padded_lstm = False: an utterance's durations do not depend on its micro-batch")
        if world == 1 and not fleet:
            if not args.no_cpu_baseline:
                from oracle.synth_weights import build_oracle
                threads = os.cpu_count() or 1
                m = build_oracle(1234)
                cpu_oracle_rate(m, w, 1, threads)
                n_cpu = min(args.cpu_baseline_utts, len(w["ids"]))
                rate, secs, dt = cpu_oracle_rate(m, w, n_cpu, threads)
                line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=threads, kind="port",
                                            sample=f"{'all' if n_cpu >= len(w['ids']) else 'first ' + str(n_cpu) + ' of the'} {len(w['ids'])} utterances of the same batch "
                                                   f"({secs:.1f} audio-s in {dt:.1f} s), oracle = fp32 PyTorch CPU restatement of the reference graph")
                del m
            if not args.no_gpu_eager:
                try:
                    line["gpu_eager_baseline"] = gpu_eager_rate(w, len(w["ids"]))
                except Exception as exc:            # reported, never fatal: it is a comparison, not the product
                    line["gpu_eager_baseline"] = dict(unavailable=f"{type(exc).__name__}: {exc}")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
