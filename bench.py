#!/usr/bin/env python
"""bench.py — audio-seconds synthesised per second (RTF^-1) of the synthesis hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path (BatchKModel.forward_batch semantics: phoneme ids + voice style rows -> 24 kHz
waveform) over one batch of synthetic utterances.  Workload at every N: BASELINE.json configs[1] per GPU — 32 utterances of
60..90 symbols (~20 Chinese characters), one voice; rank r uses seed 2+r, so per-GPU work is fixed ("weak" scaling) and
utterances are partitioned by rank with no collective on the data path.

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


def cpu_oracle_rate(m, w, n_utt, threads):
    import torch
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    o = m.forward_batch(w["ids"][:n_utt], w["ref_s"][:n_utt], w["speeds"][:n_utt], noise_seed=0)
    dt = time.perf_counter() - t0
    secs = sum(a.numel() for a in o.audio) / SR
    return secs / dt, secs, dt


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (its PyTorch graph, restated in oracle/) on the host cores."""
    if rank != 0:
        return
    import torch
    from oracle.synth_weights import build_oracle
    threads = os.cpu_count() or 1
    m = build_oracle(1234)
    w = make_workload(0)
    n_utt = args.ref_utts
    n_warm = min(max(args.warmup, 0), 1)       # a CPU step is seconds of work: one warm-up step (lazy initialisation), at most 3 timed
    for _ in range(n_warm):
        cpu_oracle_rate(m, w, 1, threads)
    rates, tot_s, tot_t = [], 0.0, 0.0
    for _ in range(max(1, min(args.steps, 3))):
        r, secs, dt = cpu_oracle_rate(m, w, n_utt, threads)
        rates.append(r); tot_s += secs; tot_t += dt
    value = tot_s / tot_t
    sample = f"first {n_utt} of the 32 utterances of configs[1] per step ({tot_s / len(rates):.1f} audio-s), fp32 PyTorch CPU eager, per-sample decode loop as kmodel.py:108-144"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(rates), warmup=n_warm, ms_per_step=1e3 * tot_t / len(rates),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="configs[1]: batch of 32 short utterances (60-90 symbols), one voice; bounded sample", sample_utterances=n_utt,
                            requested=dict(steps=args.steps, warmup=args.warmup), bounded="steps <= 3, warmup <= 1: every CPU step is seconds of work"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TTSB_BENCH_PRECISION", "f16"), choices=["f16", "bf16", "fp32"])
    ap.add_argument("--ref-utts", type=int, default=8, help="utterances per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-utts", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
    warmup = max(args.warmup, 3)
    prec = {"f16": pkg.PRECISION_FP16, "bf16": pkg.PRECISION_BF16, "fp32": pkg.PRECISION_FP32}[args.precision]

    ckpt = synthetic_checkpoint(1234)                                   # random-init weights of the Kokoro-82M architecture
    model = pkg.BatchKModel(config=synthetic_config(), model=ckpt, precision=prec, strict=True).to(f"cuda:{local}").eval()
    eng = model.engine
    w = make_workload(rank)
    ref_pinned = w["ref_s"].unsqueeze(1).contiguous().pin_memory()      # [32,1,256] like pipeline.py:603
    tokens = [len(x) for x in w["ids"]]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed steps (value) ----
    eng.profile(True)
    uid = getattr(torch.cuda.get_device_properties(local), "uuid", None)
    clocks = ClockSampler(local, uuid=("GPU-" + str(uid)) if uid is not None else None)
    clocks.open()                                          # NVML initialisation: before the warm-up, far from the timed region
    frames = None
    for _ in range(warmup):
        r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], noise=None, seed=7)
        frames = [int(a.numel()) // 600 for a in r.audio]
    audio_s = sum(frames) * 600 / SR
    sync_all()
    clocks.start()
    l0 = eng.launch_count()
    dev_ms, prof_acc = 0.0, {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.synthesize(w["ids"], w["ref_s"], w["speeds"], noise=None, seed=7)
        dev_ms += eng.last_step_ms()
        for k, v in eng.profile_read().items():
            a = prof_acc.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            for kk in a:
                a[kk] += v[kk]
    sync_all()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    eng.profile(False)

    # ---- standalone HBM-bound kernels on buffers larger than L2 (north-star: fused elementwise / iSTFT vs HBM peak) ----
    hbm_ops = {}
    if rank == 0:
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
        # length-regulation gather at a streaming size (512 utterances x 300 tokens, C = 640: 1.8 GB); `gather` above is the step's own
        # 23-29 MB, where the two launches (index map + gather) are almost all ramp
        Bg, Tg, Cg = 512, 300, 640
        dur = torch.randint(1, 7, (Bg * Tg,), generator=torch.Generator().manual_seed(3), dtype=torch.int32).cuda()
        srcg = torch.randn(Bg * Tg, Cg, device="cuda", generator=g)
        Fg = int(dur.sum())
        bg = 1e9
        for _ in range(4):
            bg = min(bg, eng.op_length_regulate(dur, [Tg * b for b in range(Bg + 1)], srcg, Fg, timed=True)[3])
        hbm_ops["gather_streaming_1p8GB"] = dict(bytes=4.0 * Cg * (Bg * Tg + Fg), ms=bg)
        del srcg, dur
    # ---- end to end through the reference-facing class (e2e) ----
    for _ in range(2):
        [a.cpu() for a in model.forward_batch(w["phonemes"], ref_pinned, w["speeds"])]
    sync_all()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        outs = [a.cpu() for a in model.forward_batch(w["phonemes"], ref_pinned, w["speeds"])]
        d2h = sum(o.numel() for o in outs) * 4 + sum(tokens) * 8
    sync_all()
    e2e_wall = time.perf_counter() - t1
    h2d = sum(tokens) * 4 + (len(tokens) + 1) * 4 + ref_pinned.numel() * 4 + len(tokens) * 4

    # ---- max over ranks, totals over ranks ----
    t = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device="cuda")
    tot = torch.tensor([audio_s, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max, e2e_max = t.tolist()
    audio_total, launches_total = tot.tolist()

    if rank == 0:
        peaks = load_peaks()
        dom = "conv_tensor" if prof_acc.get("conv_tensor", {}).get("launches", 0) else "conv_fp32"
        p = prof_acc[dom]
        ach = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0.0
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "r01j_traffic.json")
        if dom == "conv_tensor" and os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic, traffic_note = tj["dram_bytes_per_launch_mean"], tj["source"]
        roofline = dict(bound="tensor", kernel=("conv_umma_kernel (tcgen05 implicit-GEMM Conv1d)" if dom == "conv_tensor" else "conv_simt_kernel (fp32 FFMA implicit-GEMM Conv1d)"),
                        achieved=ach, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=ach / peaks["tf_sustained"], traffic=traffic, traffic_note=traffic_note,
                        peak_source=f"bf16 sustained, {peaks['source']}", launches_per_step=p["launches"] / args.steps,
                        # the same kernel on the other axis: algorithmic bytes (inputs + outputs + residuals, from the launch arguments) per second
                        hbm_achieved_gbs=(p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0),
                        hbm_frac=(p["bytes"] / (p["ms"] * 1e-3) / 1e9 / peaks["hbm"] if p["ms"] > 0 else 0.0),
                        avg_launch_ms=p["ms"] / max(p["launches"], 1), share_of_step=p["ms"] / max(dev_ms, 1e-9))
        hbm = {}
        for k in ("istft", "upsample_add", "gather"):
            q = prof_acc.get(k)
            if q and q["ms"] > 0:
                g = q["bytes"] / (q["ms"] * 1e-3) / 1e9
                hbm[k] = dict(achieved_gbs=g, frac=g / peaks["hbm"], ms_per_step=q["ms"] / args.steps)
        for k, v in hbm_ops.items():
            gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            hbm[k] = dict(achieved_gbs=gbs, frac=gbs / peaks["hbm"], ms=v["ms"], note="standalone op, buffers > L2, best of 4")
        line = dict(metric=METRIC, value=audio_total * args.steps / (dev_ms_max * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=warmup,
                    ms_per_step=dev_ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"f16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], data="synthetic",
                    config=dict(workload="configs[1]: batch of 32 short utterances (60-90 symbols each), one voice, per GPU", utterances_per_gpu=len(tokens),
                                tokens_per_gpu=sum(tokens), frames_per_gpu=sum(frames), audio_s_per_step_per_gpu=audio_s, weights="seeded random init, Kokoro-82M architecture",
                                partitioning="utterances by rank, no collective", l2="inputs larger than L2: every stage buffer of a step exceeds 126 MB",
                                text_side="fp32", frame_side=args.precision),
                    wall_ms_per_step=1e3 * wall_max / args.steps,
                    e2e=dict(value=audio_total * args.steps / e2e_max, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             api="BatchKModel.forward_batch(list[str], pinned ref_s [B,1,256], speeds) + .cpu() per waveform"),
                    gpu_launches=int(launches_total), clocks=clk, roofline=roofline, hbm_kernels=hbm,
                    algorithmic_tflops_per_s=algorithmic_flops(tokens, frames) * world / (dev_ms_max / args.steps * 1e-3) / 1e12,
                    profile_ms_per_step={k: v["ms"] / args.steps for k, v in prof_acc.items()},
                    # share of the device-timed step covered by the event-timed categories (the rest: small kernels, host gap between the halves)
                    profile_accounted_frac=sum(v["ms"] for v in prof_acc.values()) / max(dev_ms, 1e-9))
        if not args.no_cpu_baseline and world == 1:
            from oracle.synth_weights import build_oracle
            threads = os.cpu_count() or 1
            m = build_oracle(1234)
            cpu_oracle_rate(m, w, 1, threads)
            rate, secs, dt = cpu_oracle_rate(m, w, args.cpu_baseline_utts, threads)
            line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=threads, kind="port",
                                        sample=f"first {args.cpu_baseline_utts} utterances of the same batch ({secs:.1f} audio-s in {dt:.1f} s), oracle = fp32 PyTorch CPU restatement of the reference graph")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
