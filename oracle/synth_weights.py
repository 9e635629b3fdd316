"""Seeded synthetic checkpoint / voice packs / inputs in the REAL layouts.  TEST INFRASTRUCTURE.

There is no network, so ``kokoro-v1_1-zh.pth`` and ``voices/*.pt`` cannot be fetched
(SURVEY.md §8c/d).  This module builds, deterministically from a seed:
  * a checkpoint dict ``{bert, bert_encoder, predictor, decoder, text_encoder}`` of state
    dicts with ``module.``-prefixed keys and old-style ``weight_g``/``weight_v`` entries —
    the layout kokoro 0.9.4 loads (SURVEY.md §5 "Checkpoint / resume");
  * voice packs ``[510,1,256]`` fp32 (what pipeline.py:104 ``torch.load``s);
  * a vocab ``{char: id}`` with 177 symbols (ids 1..177; 0 is BOS/EOS/pad);
  * the BASELINE.json workloads as concretised in SURVEY.md §8(d).
Initialisation = PyTorch defaults, then: ``duration_proj`` rescaled/shifted so durations
average a few frames per token with spread; ``F0_proj`` rescaled/shifted so the F0 curve
mixes voiced (>10 Hz) and unvoiced spans; ``ups``/``conv_post`` ~ N(0, 0.01) as kokoro's
``init_weights`` does; Snake alphas ~ U(0.5,1.5); LayerNorm affine perturbed.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Tuple

import torch

from .kokoro_oracle import KokoroConfig, KokoroOracle


def build_oracle(seed: int = 1234) -> KokoroOracle:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        m = KokoroOracle(KokoroConfig()).eval()
        with torch.no_grad():
            # a random 12x-shared transformer layer rank-collapses (all tokens alike); damp the two residual
            # branches so token identity survives to the duration head, as it does in a trained model
            lyr = m.bert.encoder.albert_layer_groups[0].albert_layers[0]
            lyr.attention.dense.weight.mul_(0.25)
            lyr.ffn_output.weight.mul_(0.25)
            _calibrate(m)
            gen = m.decoder.generator
            for up in gen.ups:
                up.weight_v.normal_(0.0, 0.01)
                up.weight_g.copy_(up.weight_v.flatten(1).norm(dim=1).view_as(up.weight_g))
            gen.conv_post.weight_v.normal_(0.0, 0.01)
            gen.conv_post.weight_g.copy_(gen.conv_post.weight_v.flatten(1).norm(dim=1).view_as(gen.conv_post.weight_g))
            for name, prm in m.named_parameters():
                if ".alpha" in name:
                    prm.uniform_(0.5, 1.5)
                elif name.endswith("gamma") or (name.endswith("LayerNorm.weight") or name.endswith("layer_norm.weight")):
                    prm.add_(0.1 * torch.randn_like(prm))
                elif name.endswith("beta") or (name.endswith("LayerNorm.bias") or name.endswith("layer_norm.bias")):
                    prm.add_(0.1 * torch.randn_like(prm))
                elif name.endswith("weight_g"):
                    # real checkpoints have g != ||v||: exercise the fold
                    prm.mul_(1.0 + 0.2 * torch.rand_like(prm))
        torch.random.set_rng_state(g_state)
    return m


def _calibrate(m: KokoroOracle) -> None:
    """Deterministic data-dependent rescale of the two scalar heads so the synthetic model behaves like a
    trained one: token-dependent durations (mean ~4 frames/token, range 1..~15) and an F0 curve
    (mean ~100 Hz, std ~90 Hz) that crosses the 10 Hz voiced threshold."""
    g = torch.Generator().manual_seed(99)
    ids = [random_ids(g, n) for n in (50, 70, 90, 64)]
    ref = synthetic_voice_pack(7)[[49, 69, 89, 63], 0]
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
    valid = x[~mask]                                     # [sum T, 512]
    lin = p.duration_proj.linear_layer
    z = valid @ lin.weight.t()
    scale = 1.6 / z.std(dim=0).mean()
    lin.weight.mul_(scale)
    lin.bias.copy_(torch.linspace(-6.0, -2.5, 50) - scale * z.mean(dim=0))
    # F0 / N heads on one utterance's frames
    dur = torch.sigmoid(lin(x[0, : lengths[0]])).sum(-1).round().clamp(min=1).long()
    idx = torch.repeat_interleave(torch.arange(int(lengths[0])), dur)
    en = d[0, : lengths[0]].t()[None][:, :, idx]
    h, _ = p.shared(en.transpose(1, 2))
    f = h.transpose(1, 2)
    for b in p.F0:
        f = b(f, ref[:1, 128:])
    y = p.F0_proj(f)
    k = 90.0 / y.std()
    p.F0_proj.weight.mul_(k)
    p.F0_proj.bias.copy_(100.0 - k * (y.mean() - p.F0_proj.bias))
    p.N_proj.weight.mul_(4.0)


def synthetic_checkpoint(seed: int = 1234) -> Dict[str, Dict[str, torch.Tensor]]:
    return build_oracle(seed).export_checkpoint()


def synthetic_vocab() -> Dict[str, int]:
    """177 single-character symbols -> ids 1..177 (n_token = 178, id 0 reserved)."""
    chars: List[str] = []
    cp = 0x0250  # IPA extensions block onwards, plus ASCII letters: all single code points
    for ch in "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ;:,.!?-\"()' ":
        chars.append(ch)
    while len(chars) < 177:
        chars.append(chr(cp))
        cp += 1
    return {ch: i + 1 for i, ch in enumerate(chars[:177])}


def synthetic_voice_pack(seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(10_000 + seed)
    return torch.randn(510, 1, 256, generator=g) * 0.25


def random_ids(g: torch.Generator, n_symbols: int) -> List[int]:
    body = torch.randint(1, 178, (n_symbols,), generator=g).tolist()
    return [0] + body + [0]


def workload(name: str, seed: int | None = None) -> Dict:
    """BASELINE.json configs as fixed by SURVEY.md §8(d).  Returns ids (with BOS/EOS), per-utterance
    style rows [B,256], speeds."""
    if name == "single":  # config 1: '你好世界' stand-in: 14 symbols
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        ids = [random_ids(g, 14)]
        pack = synthetic_voice_pack(1)
        ref = torch.stack([pack[min(14 - 1, 509), 0]])
        return dict(ids=ids, ref_s=ref, speeds=[1.0])
    if name == "batch32":  # config 2
        g = torch.Generator().manual_seed(2 if seed is None else seed)
        lens = torch.randint(60, 91, (32,), generator=g).tolist()
        ids = [random_ids(g, n) for n in lens]
        pack = synthetic_voice_pack(1)
        ref = torch.stack([pack[min(n - 1, 509), 0] for n in lens])
        return dict(ids=ids, ref_s=ref, speeds=[1.0] * 32)
    if name == "paragraph":  # config 3: 6 sequential chunks, T 300-400
        g = torch.Generator().manual_seed(3 if seed is None else seed)
        lens = torch.randint(300, 401, (6,), generator=g).tolist()
        ids = [random_ids(g, n) for n in lens]
        pack = synthetic_voice_pack(1)
        ref = torch.stack([pack[min(n - 1, 509), 0] for n in lens])
        return dict(ids=ids, ref_s=ref, speeds=[1.0] * 6)
    if name.startswith("fleet"):  # config 5: 4096 utterances (or fleet<N>), 16 voices, speeds U(0.8,1.2)
        n = int(name[5:] or 4096)
        g = torch.Generator().manual_seed(5 if seed is None else seed)
        lens = torch.randint(40, 121, (n,), generator=g).tolist()
        ids = [random_ids(g, k) for k in lens]
        packs = [synthetic_voice_pack(100 + v) for v in range(16)]
        ref = torch.stack([packs[i % 16][min(k - 1, 509), 0] for i, k in enumerate(lens)])
        speeds = (0.8 + 0.4 * torch.rand(n, generator=g)).tolist()
        return dict(ids=ids, ref_s=ref, speeds=speeds)
    raise KeyError(name)


def vocoder_inputs(B: int, L: int, seed: int = 4) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Config 4: x [B,512,L] ~ N(0,1); f0 [B,L] 0 on ~30 % unvoiced spans else U(80,400); s [B,128]; noise."""
    g = torch.Generator().manual_seed(seed * 7919 + B * 131 + L)
    x = torch.randn(B, 512, L, generator=g)
    f0 = 80 + 320 * torch.rand(B, L, generator=g)
    span = 10
    nsp = (L + span - 1) // span
    unv = (torch.rand(B, nsp, generator=g) < 0.3).repeat_interleave(span, dim=1)[:, :L]
    f0 = torch.where(unv, torch.zeros_like(f0), f0)
    s = torch.randn(B, 128, generator=g) * 0.25
    noise = torch.randn(B, 300 * L, 9, generator=g)
    return x, s, f0, noise
