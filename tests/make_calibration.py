"""One-off: compute the three head rescalings of the synthetic checkpoint (illufly-tts_b200/synthetic_calibration.json).
Runs the CPU oracle on the UNcalibrated seeded checkpoint; deterministic.   python tests/make_calibration.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import illufly_tts_b200.synthetic as syn  # noqa: E402
from oracle.kokoro_oracle import KokoroOracle  # noqa: E402

SEED = 1234


@torch.no_grad()
def main():
    m = KokoroOracle().eval().load_checkpoint(syn.synthetic_checkpoint(SEED, calibrated=False))
    g = torch.Generator().manual_seed(99)
    ids = [syn.random_ids(g, n) for n in (50, 70, 90, 64)]
    ref = syn.synthetic_voice_pack(7)[[49, 69, 89, 63], 0]
    lengths = torch.tensor([len(x) for x in ids])
    Tm = int(lengths.max())
    padded = torch.zeros(len(ids), Tm, dtype=torch.long)
    for i, x in enumerate(ids):
        padded[i, : len(x)] = torch.tensor(x)
    mask = torch.arange(Tm)[None].expand(len(ids), -1) + 1 > lengths[:, None]
    p = m.predictor
    d_en = m.bert_encoder(m.bert(padded, (~mask).int())).transpose(-1, -2)
    d = p.text_encoder(d_en, ref[:, 128:], lengths, mask)
    x, _ = p.lstm(d)
    valid = x[~mask]
    lin = p.duration_proj.linear_layer
    z = valid @ lin.weight.t()
    scale = float(1.6 / z.std(dim=0).mean())
    bias = torch.linspace(-6.0, -2.5, 50) - scale * z.mean(dim=0)
    lin.weight.mul_(scale)
    lin.bias.copy_(bias)
    dur = torch.sigmoid(lin(x[0, : lengths[0]])).sum(-1).round().clamp(min=1).long()
    idx = torch.repeat_interleave(torch.arange(int(lengths[0])), dur)
    en = d[0, : lengths[0]].t()[None][:, :, idx]
    h, _ = p.shared(en.transpose(1, 2))
    f = h.transpose(1, 2)
    for b in p.F0:
        f = b(f, ref[:1, 128:])
    y = p.F0_proj(f)
    k = float(90.0 / y.std())
    f0_bias = float(100.0 - k * (y.mean() - p.F0_proj.bias))
    out = dict(seed=SEED, dur_scale=scale, dur_bias=[float(v) for v in bias], f0_scale=k, f0_bias=f0_bias)
    path = os.path.join(ROOT, "illufly-tts_b200", "synthetic_calibration.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path, "dur_scale", scale, "f0_scale", k, "mean dur", float(dur.float().mean()), "range", int(dur.min()), int(dur.max()))


if __name__ == "__main__":
    main()
