"""Stage-by-stage comparison of the CUDA engine with the CPU oracle (run on a GPU box):
    python tests/parity_report.py [single|pair|batch4] [fp32|bf16]
Prints max-abs / relative error per tap so a divergence can be localised in one run."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.kokoro_oracle import make_noise  # noqa: E402
from oracle.synth_weights import build_oracle, random_ids, synthetic_voice_pack, workload  # noqa: E402


def cases(name):
    if name == "single":
        return workload("single")
    g = torch.Generator().manual_seed(11)
    lens = {"pair": [20, 33], "batch4": [12, 40, 25, 40]}[name]
    ids = [random_ids(g, n) for n in lens]
    pack = synthetic_voice_pack(1)
    ref = torch.stack([pack[n - 1, 0] for n in lens])
    return dict(ids=ids, ref_s=ref, speeds=[1.0, 0.9, 1.1, 1.0][: len(lens)])


def rel(a, b):
    d = (a - b).abs().max().item()
    s = b.abs().max().item() + 1e-12
    return d, d / s


def describe(key, a, b):
    if a.shape != b.shape:
        print(f"{key:12s} SHAPE engine {tuple(a.shape)} oracle {tuple(b.shape)}")
        return
    e = (a - b).abs()
    d, rr = rel(a, b)
    worst = e.flatten().argmax().item()
    rms = (e.double() ** 2).mean().sqrt().item()
    frac = (e > 1e-3 * (b.abs().max().item() + 1e-12)).double().mean().item()
    nz = int((e > 0).sum())
    print(f"{key:12s} shape {tuple(a.shape)} max-abs {d:.3e} rel {rr:.3e} rms {rms:.3e} frac(>1e-3 rel) {frac:.2e} n_differ {nz} "
          f"(worst at row {worst // a.shape[1]}, col {worst % a.shape[1]}) nan={bool(torch.isnan(a).any())}")


def vocoder_report(m, eng, o, w):
    """generator only, fed with the oracle's own gen_in / F0 (bit-identical inputs)"""
    B = len(w["ids"])
    Fs = [int(p.sum()) for p in o.pred_dur]
    Lmax = 2 * max(Fs)
    x = torch.zeros(B, 512, Lmax)
    f0 = torch.zeros(B, Lmax)
    nz = torch.zeros(B, 300 * Lmax, 9)
    for i in range(B):
        L = 2 * Fs[i]
        x[i, :, :L] = o.taps[i]["gen_in"][0]
        f0[i, :L] = o.taps[i]["F0"]
        nz[i, : 300 * L] = make_noise(0, i, Fs[i])
    s = w["ref_s"][:, :128]
    audio = eng.vocoder(x, s, f0, nz, lengths=[2 * f for f in Fs])
    print("---- vocoder only, oracle inputs ----")
    for key in ["har_source", "har", "gen_stage0", "gen_stage1", "conv_post"]:
        a = eng.read_tap(key)
        parts = []
        for i in range(B):
            v = o.taps[i][key]
            v = v[0] if v.dim() == 3 else v
            v = v[:, None] if v.dim() == 1 else v.t()
            parts.append(v)
        describe(key, a, torch.cat(parts, 0))
    har_e = eng.read_tap("har")
    row = 0
    for i in range(B):
        a, b = audio[i].cpu(), o.audio[i]
        err = a - b
        snr = 10 * torch.log10((b ** 2).sum() / (err ** 2).sum().clamp_min(1e-30)).item()
        prof = [f"{err[k * len(err) // 8:(k + 1) * len(err) // 8].abs().max().item():.1e}" for k in range(8)]
        print(f"audio {i}: n={a.numel()} max-abs {err.abs().max().item():.3e} snr {snr:.1f} dB profile {prof}")
        # oracle with the degenerate STFT frame 0 pinned to the engine's phases
        L = 2 * Fs[i]
        t2 = {}
        b2 = m.vocoder(x[i:i + 1, :, :L], s[i:i + 1], f0[i:i + 1, :L], nz[i:i + 1, :300 * L], t2, har_e[row:row + 60 * L + 1].t()[None].contiguous())[0]
        err = a - b2
        prof = [f"{err[k * len(err) // 8:(k + 1) * len(err) // 8].abs().max().item():.1e}" for k in range(8)]
        print(f"   vs oracle fed the engine source spectrum: max-abs {err.abs().max().item():.3e} profile {prof}")
        nrow = 60 * L + 1
        for key in ["gen_stage0", "gen_stage1", "conv_post"]:
            aa = eng.read_tap(key)
            r0 = {"gen_stage0": (row - i) // 6, "gen_stage1": row, "conv_post": row}[key]
            n = {"gen_stage0": 10 * L, "gen_stage1": nrow, "conv_post": nrow}[key]
            describe("  " + key, aa[r0:r0 + n], t2[key][0].t())
        row += nrow


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "single"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    import illufly_tts_b200 as pkg
    m = build_oracle(1234)
    w = cases(name)
    t = time.time()
    o = m.forward_batch(w["ids"], w["ref_s"], w["speeds"], noise_seed=0, want_taps=True)
    print(f"oracle: {time.time() - t:.2f}s  frames={[int(p.sum()) for p in o.pred_dur]}")
    eng = pkg.Engine(m.export_checkpoint(), 0, {"fp32": pkg.PRECISION_FP32, "bf16": pkg.PRECISION_BF16, "f16": pkg.PRECISION_FP16}[prec])
    eng.enable_taps(True)
    B = len(w["ids"])
    # durations first (no noise needed), then the full run with the oracle's noise
    noise = [make_noise(0, i, int(o.pred_dur[i].sum())) for i in range(B)]
    forced = [p.tolist() for p in o.pred_dur]
    t = time.time()
    r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], noise=noise, want_indices=True, forced_dur=forced)
    torch.cuda.synchronize()
    print(f"engine: {time.time() - t:.3f}s  launches={eng.launch_count()}")
    ok = True
    for i in range(B):
        same = torch.equal(r.pred_dur[i], o.pred_dur[i])
        margin = (o.duration[i] - o.duration[i].floor() - 0.5).abs().min().item()
        dd = (r.duration[i] - o.duration[i]).abs().max().item()
        print(f"utt {i}: pred_dur equal={same} (|dur diff| max {dd:.2e}, oracle margin to .5 {margin:.3e}); indices equal={torch.equal(r.indices[i], o.indices[i])}")
        ok &= same
    # taps: engine taps are ragged-concatenated channels-last; oracle taps are per utterance channels-first
    Ts = [len(x) for x in w["ids"]]
    Fs = [int(p.sum()) for p in o.pred_dur]

    def cat_oracle(key, cl=True):
        parts = []
        for i in range(B):
            v = o.taps[i][key]
            if v.dim() == 3:
                v = v[0]
            if v.dim() == 1:
                v = v[:, None]
            elif cl:
                v = v.t()
            parts.append(v)
        return torch.cat(parts, 0)

    for key, okey, cl in [("d", "d", False), ("t_en", "t_en", True), ("en", "en", True), ("asr", "asr", True), ("F0", "F0", True), ("N", "N", True),
                          ("gen_in", "gen_in", True), ("har_source", "har_source", True), ("har", "har", True), ("gen_stage0", "gen_stage0", True),
                          ("gen_stage1", "gen_stage1", True), ("conv_post", "conv_post", True)]:
        try:
            a = eng.read_tap(key)
        except Exception as exc:
            print(f"{key:12s} <no tap: {exc}>")
            continue
        describe(key, a, cat_oracle(okey, cl))
    for i in range(B):
        a, b = r.audio[i].cpu(), o.audio[i]
        if a.shape != b.shape:
            print(f"audio {i}: SHAPE {tuple(a.shape)} vs {tuple(b.shape)}")
            continue
        err = (a - b)
        snr = 10 * torch.log10((b ** 2).sum() / (err ** 2).sum().clamp_min(1e-30)).item()
        print(f"audio {i}: n={a.numel()} max-abs {err.abs().max().item():.3e} snr {snr:.1f} dB  (|ref| max {b.abs().max().item():.3f})")
    vocoder_report(m, eng, o, w)
    # predicted (not forced) durations through the whole path
    r2 = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], noise=None, seed=1)
    print("unforced run frames:", [int(a.numel()) // 600 for a in r2.audio], "oracle:", Fs, "tokens:", Ts)


if __name__ == "__main__":
    main()
