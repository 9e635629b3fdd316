"""Seeded synthetic checkpoint / voice packs / workloads in the REAL layouts.

There is no network, so ``kokoro-v1_1-zh.pth`` and ``voices/*.pt`` cannot be fetched.  This module
builds, deterministically from a seed and WITHOUT running any model code:
  * a checkpoint ``{bert, bert_encoder, predictor, decoder, text_encoder}`` of state dicts with ``module.``-prefixed keys
    and old-style ``weight_g``/``weight_v`` entries — the layout kokoro 0.9.4 loads.  Names and shapes come
    from ``checkpoint_manifest.json`` (dumped from the architecture: 548 tensors, 81.8 M parameters);
  * voice packs ``[510,1,256]`` fp32 (what pipeline.py:104 ``torch.load``s);
  * a vocab ``{char: id}`` with 177 symbols (ids 1..177; 0 is BOS/EOS/pad);
  * the BASELINE.json workloads as concretised in SURVEY.md §8(d).
Values: PyTorch-default-like initialisation by tensor kind; ``ups``/``conv_post`` ~ N(0, 0.01) as kokoro's ``init_weights``
does; Snake alphas ~ U(0.5,1.5); LayerNorm affine perturbed; weight_g != ||v||.  Three scalar heads are then rescaled with
constants from ``synthetic_calibration.json`` (computed once by tests/make_calibration.py) so the random model behaves like
a trained one: token-dependent durations of a few frames per token, an F0 curve that crosses the 10 Hz voiced threshold.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_MANIFEST = os.path.join(_HERE, "checkpoint_manifest.json")
_CALIB = os.path.join(_HERE, "synthetic_calibration.json")


def manifest() -> Dict[str, Dict[str, List[int]]]:
    with open(_MANIFEST) as f:
        return json.load(f)


def _init(key: str, shape: List[int], g: torch.Generator, made: Dict[str, torch.Tensor]) -> torch.Tensor:
    def U(bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def N(std, mean=0.0):
        return torch.randn(shape, generator=g) * std + mean
    leaf = key.rsplit(".", 1)[-1]
    if ".alpha" in key:
        return torch.rand(shape, generator=g) + 0.5
    if leaf in ("gamma",) or key.endswith("LayerNorm.weight") or key.endswith("layer_norm.weight"):
        return N(0.1, 1.0)
    if leaf in ("beta",) or key.endswith("LayerNorm.bias") or key.endswith("layer_norm.bias"):
        return N(0.1)
    if "embedding" in key and leaf == "weight" and len(shape) == 2 and "mapping" not in key:
        return N(1.0)
    if "lstm" in key or ".shared." in key:
        return U(1.0 / 16.0)
    if leaf == "weight_g":
        v = made[key.split(".", 1)[1][:-1] + "v"]
        return v.flatten(1).norm(dim=1).view(shape) * (1.0 + 0.2 * torch.rand(shape, generator=g))
    if leaf == "weight_v":
        if ".ups." in key or "conv_post" in key:
            return N(0.01)
        fan_in = shape[1] * (shape[2] if len(shape) > 2 else 1)
        return U(1.0 / math.sqrt(fan_in))
    if leaf == "weight":
        fan_in = shape[1] * (shape[2] if len(shape) > 2 else 1) if len(shape) > 1 else shape[0]
        w = U(1.0 / math.sqrt(fan_in))
        # a random 12x-shared transformer layer rank-collapses (all tokens alike): damp the two residual branches so
        # token identity survives to the duration head, as it does in a trained model
        if key.endswith("attention.dense.weight") or key.endswith("ffn_output.weight"):
            w = w * 0.25
        return w
    if leaf == "bias":
        return U(0.05)
    raise KeyError(f"no init rule for {key}")


def synthetic_checkpoint(seed: int = 1234, calibrated: bool = True) -> Dict[str, Dict[str, torch.Tensor]]:
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for mod, entries in manifest().items():
        made: Dict[str, torch.Tensor] = {}
        keys = sorted(entries, key=lambda k: (k.endswith("weight_g"), k))     # weight_v before its weight_g
        for k in keys:
            made[k] = _init(f"{mod}.{k}", entries[k], g, made).to(torch.float32).contiguous()
        out[mod] = {"module." + k: made[k] for k in entries}
    if calibrated:
        if not os.path.exists(_CALIB):
            raise FileNotFoundError(f"{_CALIB} missing: run tests/make_calibration.py")
        with open(_CALIB) as f:
            c = json.load(f)
        if c["seed"] != seed:
            raise ValueError(f"synthetic_calibration.json was computed for seed {c['seed']}, not {seed}")
        p = out["predictor"]
        p["module.duration_proj.linear_layer.weight"] = p["module.duration_proj.linear_layer.weight"] * c["dur_scale"]
        p["module.duration_proj.linear_layer.bias"] = torch.tensor(c["dur_bias"], dtype=torch.float32)
        p["module.F0_proj.weight"] = p["module.F0_proj.weight"] * c["f0_scale"]
        p["module.F0_proj.bias"] = torch.tensor([c["f0_bias"]], dtype=torch.float32)
        p["module.N_proj.weight"] = p["module.N_proj.weight"] * 4.0
    return out


def synthetic_vocab() -> Dict[str, int]:
    """177 single-character symbols -> ids 1..177 (n_token = 178, id 0 reserved)."""
    chars: List[str] = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ;:,.!?-\"()' ")
    cp = 0x0250  # IPA extensions block onwards: all single code points
    while len(chars) < 177:
        chars.append(chr(cp))
        cp += 1
    return {ch: i + 1 for i, ch in enumerate(chars[:177])}


def synthetic_config() -> Dict:
    """The fields of the HF config.json that the engine's host side reads."""
    return {"vocab": synthetic_vocab(), "n_token": 178, "hidden_dim": 512, "style_dim": 128, "n_layer": 3, "max_dur": 50,
            "plbert": {"hidden_size": 768, "num_attention_heads": 12, "intermediate_size": 2048, "max_position_embeddings": 512,
                       "num_hidden_layers": 12}}


def synthetic_voice_pack(seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(10_000 + seed)
    return torch.randn(510, 1, 256, generator=g) * 0.25


def random_ids(g: torch.Generator, n_symbols: int) -> List[int]:
    body = torch.randint(1, 178, (n_symbols,), generator=g).tolist()
    return [0] + body + [0]


def ids_to_phonemes(ids: List[int]) -> str:
    inv = {v: k for k, v in synthetic_vocab().items()}
    return "".join(inv[i] for i in ids[1:-1])


def workload(name: str, seed: int | None = None) -> Dict:
    """BASELINE.json configs as fixed by SURVEY.md §8(d): ids (with BOS/EOS), per-utterance style rows [B,256], speeds."""
    def pack_rows(pack, lens):
        return torch.stack([pack[min(n - 1, 509), 0] for n in lens])      # pipeline.py:596-601
    if name == "single":      # config 1: stand-in for the 14 symbols of the 4-character greeting
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        ids = [random_ids(g, 14)]
        return dict(ids=ids, ref_s=pack_rows(synthetic_voice_pack(1), [14]), speeds=[1.0])
    if name == "batch32":     # config 2
        g = torch.Generator().manual_seed(2 if seed is None else seed)
        lens = torch.randint(60, 91, (32,), generator=g).tolist()
        return dict(ids=[random_ids(g, n) for n in lens], ref_s=pack_rows(synthetic_voice_pack(1), lens), speeds=[1.0] * 32)
    if name == "paragraph":   # config 3: 6 sequential chunks
        g = torch.Generator().manual_seed(3 if seed is None else seed)
        lens = torch.randint(300, 401, (6,), generator=g).tolist()
        return dict(ids=[random_ids(g, n) for n in lens], ref_s=pack_rows(synthetic_voice_pack(1), lens), speeds=[1.0] * 6)
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
    """Config 4: x [B,512,L] ~ N(0,1); f0 [B,L] 0 on ~30 % unvoiced spans else U(80,400); s [B,128]; noise [B,300L,9]."""
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
