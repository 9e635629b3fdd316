"""Parity protocol shared by the GPU tests and __graft_entry__.smoke()  (DESIGN.md "Parity protocol").

The synthesis graph has two points where the REFERENCE ITSELF is not reproducible to 1e-3, on any hardware:
  1. F0 -> harmonic phase is an integrator (phase = 2 pi 300 cumsum(f0 h / 24000)): a 1e-6 relative change of F0 (fp32
     rounding of a different summation order in the F0 predictor) moves the waveform by > 1e-2 — the CPU oracle fails a
     1e-3 comparison against itself under that perturbation (tests/test_oracle.py::test_reference_is_chaotic_in_f0).
  2. angle() of the source STFT is discontinuous: where Im X ~ 0 and Re X < 0 it is +-pi by the SIGN of rounding noise —
     always at frame 0 (reflect padding makes it exactly symmetric) and wherever a bin crosses the negative real axis
     (a few % of utterances have such a crossing within 1e-7).  The sign depends on the FFT library.
So parity is checked in two halves that meet at the generator input:
  front half : ids -> pred_dur / indices bit-exact; duration, d, t_en, en, asr, F0, N, gen_in within 1e-4 relative;
  back half  : (1) the engine's source spectrum [mag | angle] equals the oracle's on the same F0 as COMPLEX numbers (1e-5);
               (2) the oracle's generator fed with bit-identical inputs (the engine's gen_in / F0 / [mag | angle] taps):
               waveform max-abs <= 1e-3 (fp32) or SNR >= 40 dB (tensor-core modes)."""
from __future__ import annotations

from typing import Dict, List

import torch


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    err = (test.double() - ref.double())
    return float(10 * torch.log10((ref.double() ** 2).sum() / (err ** 2).sum().clamp_min(1e-300)))


def rel_err(test: torch.Tensor, ref: torch.Tensor) -> float:
    return float((test - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def oracle_taps_cat(o, key: str) -> torch.Tensor:
    """oracle taps are per utterance, channels-first; the engine's are ragged-concatenated channels-last"""
    parts = []
    for t in o.taps:
        v = t[key]
        v = v[0] if v.dim() == 3 else v
        if key == "d":
            parts.append(v)
        else:
            parts.append(v[:, None] if v.dim() == 1 else v.t())
    return torch.cat(parts, 0)


def front_half(eng, oracle, w, padded_lstm: bool = True, frame_tol: float = 1e-4, o=None) -> Dict:
    """runs oracle + engine on workload w, checks the front half, returns everything the back half needs
    (``o``: an oracle run of the same workload made earlier - a batch-32 run is ~15 s of CPU and serves every precision mode)"""
    from oracle.kokoro_oracle import make_noise
    if o is None:
        o = oracle.forward_batch(w["ids"], w["ref_s"], w["speeds"], noise_seed=0, want_taps=True)
    B = len(w["ids"])
    Fs = [int(p.sum()) for p in o.pred_dur]
    noise = [make_noise(0, i, Fs[i]) for i in range(B)]
    eng.enable_taps(True)
    try:
        r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], noise=noise, want_indices=True, padded_lstm=padded_lstm)
        taps = {k: eng.read_tap(k) for k in ("d", "t_en", "en", "asr", "F0", "N", "gen_in", "har")}
    finally:
        eng.enable_taps(False)
    report = {}
    for i in range(B):
        assert torch.equal(r.pred_dur[i], o.pred_dur[i]), f"utt {i}: predicted durations differ from the oracle"
        assert torch.equal(r.indices[i], o.indices[i]), f"utt {i}: alignment index map differs from the oracle"
        assert r.audio[i].numel() == 600 * Fs[i]
        assert float((r.duration[i] - o.duration[i]).abs().max()) < 1e-4
    report["duration_abs"] = max(float((r.duration[i] - o.duration[i]).abs().max()) for i in range(B))
    report["duration_margin"] = min(float((o.duration[i] - o.duration[i].floor() - 0.5).abs().min()) for i in range(B))
    for k in ("d", "t_en", "en", "asr", "F0", "N", "gen_in"):
        report[k] = rel_err(taps[k], oracle_taps_cat(o, k))
        # the text side is fp32-class in every precision mode; frame_tol: one bar for the frame-side taps, or one per tap
        tol = (frame_tol.get(k, 1e-4) if isinstance(frame_tol, dict) else frame_tol) if k in ("F0", "N", "gen_in") else 1e-4
        assert report[k] < tol, f"tap {k}: relative error {report[k]:.3e} >= {tol}"
    return dict(o=o, r=r, taps=taps, noise=noise, Fs=Fs, report=report)


def e2e_records(fh) -> List[Dict]:
    """END TO END: the engine's waveform against the oracle's own waveform for the same ids / style / speeds / injected noise -
    no tap is substituted anywhere.  This is the number the north star names; it is reported, and bounded only loosely, because
    the reference graph itself is chaotic in F0 (tests/test_oracle.py::test_reference_is_chaotic_in_f0: the oracle against
    itself under a 1e-6 relative F0 perturbation gives ~19 dB)."""
    out = []
    for i in range(len(fh["Fs"])):
        a, b = fh["r"].audio[i].cpu(), fh["o"].audio[i]
        out.append(dict(max_abs=float((a - b).abs().max()), snr=snr_db(b, a), n=a.numel(), ref_peak=float(b.abs().max())))
    return out


def summarise(recs: List[Dict]) -> Dict:
    """worst case over the utterances of a workload"""
    return dict(utterances=len(recs), samples=sum(r["n"] for r in recs), max_abs_worst=max(r["max_abs"] for r in recs),
                snr_min=min(r["snr"] for r in recs), snr_median=sorted(r["snr"] for r in recs)[len(recs) // 2])


def record(case: str, mode: str, **fields):
    """append one line to the parity report (tests/make_parity_table.py turns it into profiles/rNN_parity.md).  The report goes
    to $TTSB_PARITY_REPORT, else to gpurun_out/parity_report.jsonl when that directory exists, else nowhere."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.environ.get("TTSB_PARITY_REPORT")
    if not path and os.path.isdir(os.path.join(root, "gpurun_out")):
        path = os.path.join(root, "gpurun_out", "parity_report.jsonl")
    if not path:
        return
    with open(path, "a") as f:
        f.write(json.dumps(dict(case=case, mode=mode, **fields)) + "\n")


def spectrum_error(har_a: torch.Tensor, har_b: torch.Tensor) -> float:
    """[.., 22, frames] features [mag | angle] compared as complex numbers (robust to +-pi wraps)"""
    za = torch.polar(har_a[..., :11, :].double(), har_a[..., 11:, :].double())
    zb = torch.polar(har_b[..., :11, :].double(), har_b[..., 11:, :].double())
    return float((za - zb).abs().max())


def vocoder_check(oracle, x, s, f0, noise, har_engine, audio_engine, tol_abs=None, min_snr=None, who="") -> Dict:
    """x [1,512,L], s [1,128], f0 [1,L], noise [1,300L,9], har_engine [60L+1, 22] (engine tap rows of this utterance)"""
    har_e = har_engine.t()[None].contiguous()
    taps = {}
    ref = oracle.vocoder(x, s, f0, noise, taps, har_e)[0]
    spec = spectrum_error(har_e, taps["har_own"])
    assert spec < 1e-5, f"{who}: source spectrum differs from the oracle's by {spec:.3e}"
    a = audio_engine.cpu()
    rec = dict(max_abs=float((a - ref).abs().max()), snr=snr_db(ref, a), n=a.numel(), spectrum_err=spec)
    if tol_abs is not None:
        assert rec["max_abs"] <= tol_abs, f"{who}: waveform max-abs {rec['max_abs']:.3e} > {tol_abs}"
    if min_snr is not None:
        assert rec["snr"] >= min_snr, f"{who}: waveform SNR {rec['snr']:.1f} dB < {min_snr}"
    return rec


def back_half(eng, oracle, w, fh, tol_abs=None, min_snr=None) -> List[Dict]:
    """oracle generator on the ENGINE's gen_in / F0 / source-spectrum taps (bit-identical inputs) vs the engine's waveform"""
    out = []
    rowL, rowS = 0, 0
    for i, F in enumerate(fh["Fs"]):
        L = 2 * F
        x = fh["taps"]["gen_in"][rowL:rowL + L].t()[None].contiguous()
        f0 = fh["taps"]["F0"][rowL:rowL + L, 0][None].contiguous()
        har = fh["taps"]["har"][rowS:rowS + 60 * L + 1]
        rec = vocoder_check(oracle, x, w["ref_s"][i:i + 1, :128], f0, fh["noise"][i][None], har, fh["r"].audio[i], None, None, f"utt {i}")
        if tol_abs is not None:
            assert rec["max_abs"] <= tol_abs, f"utt {i}: waveform max-abs {rec['max_abs']:.3e} > {tol_abs}"
        if min_snr is not None:
            assert rec["snr"] >= min_snr, f"utt {i}: waveform SNR {rec['snr']:.1f} dB < {min_snr}"
        out.append(rec)
        rowL += L
        rowS += 60 * L + 1
    return out
