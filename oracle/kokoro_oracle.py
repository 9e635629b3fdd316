"""CPU oracle for the illufly-tts synthesis hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``illufly-tts_b200``) never does.

What this is
------------
A plain fp32 PyTorch-CPU restatement of the graph that the reference runs behind
``BatchKModel`` (/root/reference/src/illufly_tts/core/kmodel.py:5-150).  The reference
itself holds none of the arithmetic: ``kmodel.py:1`` subclasses ``kokoro.model.KModel``
from the third-party PyPI package ``kokoro==0.9.4`` (pinned at
/root/reference/poetry.lock:939-955), which is not vendored, not installed in this
image and not fetchable (no network).  The module maths below therefore restates the
published kokoro 0.9.4 sources (``kokoro/model.py``, ``kokoro/modules.py``,
``kokoro/istftnet.py``) and the wiring follows the reference's
own call sites line by line (each function cites them).

PARITY UNPINNED: the reference's tests hold no golden vectors, known-answer tests or
tolerances for this path and neither ``kokoro`` nor its weights
can be run here, so this oracle is pinned only structurally:
  * the per-module parameter split equals the published Kokoro-82M split
    (tests/test_oracle.py::test_param_counts),
  * state-dict key names equal the real checkpoint layout (``weight_g``/``weight_v``,
    ``module.`` prefix handled) so ``kokoro-v1_1-zh.pth`` would load,
  * the ALBERT restatement is bit-compared with ``transformers.AlbertModel``
    (available in this image), the LSTM/STFT/iSTFT/interpolate pieces are torch's own,
  * geometry: audio length == 600 * sum(pred_dur), STFT frames == 120 F + 1.

Randomness: kokoro's ``SineGen`` draws from the unseeded global RNG; parity is only
definable with the noise injected, so every entry point takes ``noise`` explicitly
(SURVEY.md §7 hard part 1, Appendix C-4: ``rand_ini`` has no effect and is omitted).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import weight_norm  # old-style: weight_g / weight_v keys, as in the checkpoint


# --------------------------------------------------------------------------------------
# config (config.json of hexgrad/Kokoro-82M-v1.1-zh)
# --------------------------------------------------------------------------------------
@dataclass
class KokoroConfig:
    n_token: int = 178
    hidden_dim: int = 512
    style_dim: int = 128
    n_layer: int = 3
    max_dur: int = 50
    text_encoder_kernel_size: int = 5
    # plbert
    bert_hidden: int = 768
    bert_heads: int = 12
    bert_inter: int = 2048
    bert_max_pos: int = 512
    bert_layers: int = 12
    bert_emb: int = 128
    # istftnet
    upsample_rates: tuple = (10, 6)
    upsample_kernel_sizes: tuple = (20, 12)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    gen_istft_n_fft: int = 20
    gen_istft_hop_size: int = 5
    sampling_rate: int = 24000


# --------------------------------------------------------------------------------------
# A.1 ALBERT (transformers.AlbertModel restated; key names identical)
# --------------------------------------------------------------------------------------
def gelu_new(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class _AlbertEmbeddings(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.n_token, c.bert_emb, padding_idx=0)
        self.position_embeddings = nn.Embedding(c.bert_max_pos, c.bert_emb)
        self.token_type_embeddings = nn.Embedding(2, c.bert_emb)
        self.LayerNorm = nn.LayerNorm(c.bert_emb, eps=1e-12)

    def forward(self, ids):
        T = ids.shape[1]
        pos = torch.arange(T, device=ids.device)[None]
        x = self.word_embeddings(ids) + self.token_type_embeddings(torch.zeros_like(ids))
        x = x + self.position_embeddings(pos)
        return self.LayerNorm(x)


class _AlbertAttention(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        self.h = c.bert_heads
        self.dh = c.bert_hidden // c.bert_heads
        self.query = nn.Linear(c.bert_hidden, c.bert_hidden)
        self.key = nn.Linear(c.bert_hidden, c.bert_hidden)
        self.value = nn.Linear(c.bert_hidden, c.bert_hidden)
        self.dense = nn.Linear(c.bert_hidden, c.bert_hidden)
        self.LayerNorm = nn.LayerNorm(c.bert_hidden, eps=1e-12)

    def forward(self, x, add_mask):
        B, T, H = x.shape
        q = self.query(x).view(B, T, self.h, self.dh).transpose(1, 2)
        k = self.key(x).view(B, T, self.h, self.dh).transpose(1, 2)
        v = self.value(x).view(B, T, self.h, self.dh).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / math.sqrt(self.dh) + add_mask
        p = torch.softmax(s, dim=-1)
        ctx = (p @ v).transpose(1, 2).reshape(B, T, H)
        return self.LayerNorm(x + self.dense(ctx))


class _AlbertLayer(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        self.full_layer_layer_norm = nn.LayerNorm(c.bert_hidden, eps=1e-12)
        self.attention = _AlbertAttention(c)
        self.ffn = nn.Linear(c.bert_hidden, c.bert_inter)
        self.ffn_output = nn.Linear(c.bert_inter, c.bert_hidden)

    def forward(self, x, add_mask):
        a = self.attention(x, add_mask)
        f = self.ffn_output(gelu_new(self.ffn(a)))
        return self.full_layer_layer_norm(f + a)


class _AlbertLayerGroup(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.albert_layers = nn.ModuleList([_AlbertLayer(c)])


class _AlbertEncoder(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        self.n = c.bert_layers
        self.embedding_hidden_mapping_in = nn.Linear(c.bert_emb, c.bert_hidden)
        self.albert_layer_groups = nn.ModuleList([_AlbertLayerGroup(c)])

    def forward(self, x, add_mask):
        x = self.embedding_hidden_mapping_in(x)
        layer = self.albert_layer_groups[0].albert_layers[0]
        for _ in range(self.n):  # 12 applications of the ONE shared layer
            x = layer(x, add_mask)
        return x


class Albert(nn.Module):
    """kokoro ``CustomAlbert``: AlbertModel returning last_hidden_state.
    Called at kmodel.py:70 as ``self.bert(padded_ids, attention_mask=(~text_mask).int())``."""

    def __init__(self, c: KokoroConfig):
        super().__init__()
        self.embeddings = _AlbertEmbeddings(c)
        self.encoder = _AlbertEncoder(c)
        self.pooler = nn.Linear(c.bert_hidden, c.bert_hidden)  # unused by the path, present in the checkpoint

    def forward(self, ids, attention_mask):
        add = (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * torch.finfo(torch.float32).min
        return self.encoder(self.embeddings(ids), add)


# --------------------------------------------------------------------------------------
# A.3/A.4/A.6-A.8 prosody predictor
# --------------------------------------------------------------------------------------
class LinearNorm(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.linear_layer = nn.Linear(i, o)

    def forward(self, x):
        return self.linear_layer(x)


class AdaLayerNorm(nn.Module):
    def __init__(self, style_dim, channels, eps=1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.fc = nn.Linear(style_dim, channels * 2)

    def forward(self, x, s):  # x [B,T,C]
        h = self.fc(s)[:, None, :]
        gamma, beta = h[..., : self.channels], h[..., self.channels:]
        x = F.layer_norm(x, (self.channels,), eps=self.eps)
        return (1 + gamma) * x + beta


class DurationEncoder(nn.Module):
    """A.3 — called at kmodel.py:94."""

    def __init__(self, sty_dim, d_model, nlayers):
        super().__init__()
        self.lstms = nn.ModuleList()
        for _ in range(nlayers):
            self.lstms.append(nn.LSTM(d_model + sty_dim, d_model // 2, 1, batch_first=True, bidirectional=True))
            self.lstms.append(AdaLayerNorm(sty_dim, d_model))

    def forward(self, x, style, text_lengths, m):  # x [B,C,T], m [B,T] True = pad
        Tm = m.shape[-1]
        x = x.transpose(1, 2)  # [B,T,C]
        s = style[:, None, :].expand(-1, Tm, -1)
        x = torch.cat([x, s], dim=-1).masked_fill(m[..., None], 0.0)
        for block in self.lstms:
            if isinstance(block, AdaLayerNorm):
                x = block(x, style)
                x = torch.cat([x, s], dim=-1).masked_fill(m[..., None], 0.0)
            else:
                packed = nn.utils.rnn.pack_padded_sequence(x, text_lengths.cpu(), batch_first=True, enforce_sorted=False)
                y, _ = block(packed)
                y, _ = nn.utils.rnn.pad_packed_sequence(y, batch_first=True)
                x = torch.zeros(y.shape[0], Tm, y.shape[2], device=y.device)
                x[:, : y.shape[1]] = y
        return x  # [B,T,640]


class AdaIN1d(nn.Module):
    """A.8 — InstanceNorm1d (biased var, eps 1e-5; affine weight=1/bias=0 absent from the checkpoint)."""

    def __init__(self, style_dim, num_features):
        super().__init__()
        self.norm = nn.InstanceNorm1d(num_features, affine=True)
        self.fc = nn.Linear(style_dim, num_features * 2)

    def forward(self, x, s):
        h = self.fc(s)[:, :, None]
        gamma, beta = torch.chunk(h, 2, dim=1)
        return (1 + gamma) * self.norm(x) + beta


class AdainResBlk1d(nn.Module):
    """A.7"""

    def __init__(self, dim_in, dim_out, style_dim, upsample=False):
        super().__init__()
        self.up = upsample
        self.learned_sc = dim_in != dim_out
        self.conv1 = weight_norm(nn.Conv1d(dim_in, dim_out, 3, 1, 1))
        self.conv2 = weight_norm(nn.Conv1d(dim_out, dim_out, 3, 1, 1))
        self.norm1 = AdaIN1d(style_dim, dim_in)
        self.norm2 = AdaIN1d(style_dim, dim_out)
        if self.learned_sc:
            self.conv1x1 = weight_norm(nn.Conv1d(dim_in, dim_out, 1, 1, 0, bias=False))
        if upsample:
            self.pool = weight_norm(nn.ConvTranspose1d(dim_in, dim_in, 3, stride=2, groups=dim_in, padding=1, output_padding=1))
        else:
            self.pool = nn.Identity()

    def forward(self, x, s):
        r = F.leaky_relu(self.norm1(x, s), 0.2)
        r = self.pool(r)
        r = self.conv1(r)
        r = F.leaky_relu(self.norm2(r, s), 0.2)
        r = self.conv2(r)
        sc = F.interpolate(x, scale_factor=2, mode="nearest") if self.up else x
        if self.learned_sc:
            sc = self.conv1x1(sc)
        return (r + sc) * torch.rsqrt(torch.tensor(2.0))


class ProsodyPredictor(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        d, sd = c.hidden_dim, c.style_dim
        self.text_encoder = DurationEncoder(sd, d, c.n_layer)
        self.lstm = nn.LSTM(d + sd, d // 2, 1, batch_first=True, bidirectional=True)
        self.duration_proj = LinearNorm(d, c.max_dur)
        self.shared = nn.LSTM(d + sd, d // 2, 1, batch_first=True, bidirectional=True)
        self.F0 = nn.ModuleList([AdainResBlk1d(d, d, sd), AdainResBlk1d(d, d // 2, sd, upsample=True), AdainResBlk1d(d // 2, d // 2, sd)])
        self.N = nn.ModuleList([AdainResBlk1d(d, d, sd), AdainResBlk1d(d, d // 2, sd, upsample=True), AdainResBlk1d(d // 2, d // 2, sd)])
        self.F0_proj = nn.Conv1d(d // 2, 1, 1, 1, 0)
        self.N_proj = nn.Conv1d(d // 2, 1, 1, 1, 0)

    def F0Ntrain(self, x, s):  # x [B,640,F] — kmodel.py:133
        x, _ = self.shared(x.transpose(1, 2))
        x = x.transpose(1, 2)
        f0 = x
        for b in self.F0:
            f0 = b(f0, s)
        f0 = self.F0_proj(f0)
        n = x
        for b in self.N:
            n = b(n, s)
        n = self.N_proj(n)
        return f0.squeeze(1), n.squeeze(1)


# --------------------------------------------------------------------------------------
# A.9 text encoder
# --------------------------------------------------------------------------------------
class ChannelLayerNorm(nn.Module):
    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))

    def forward(self, x):  # [B,C,T]
        return F.layer_norm(x.transpose(1, -1), (self.channels,), self.gamma, self.beta, self.eps).transpose(1, -1)


class TextEncoder(nn.Module):
    """A.9 — called per sample at kmodel.py:138."""

    def __init__(self, c: KokoroConfig):
        super().__init__()
        ch, k = c.hidden_dim, c.text_encoder_kernel_size
        self.embedding = nn.Embedding(c.n_token, ch)
        self.cnn = nn.ModuleList()
        for _ in range(c.n_layer):
            self.cnn.append(nn.Sequential(weight_norm(nn.Conv1d(ch, ch, k, padding=(k - 1) // 2)), ChannelLayerNorm(ch), nn.LeakyReLU(0.2), nn.Dropout(0.2)))
        self.lstm = nn.LSTM(ch, ch // 2, 1, batch_first=True, bidirectional=True)

    def forward(self, ids, lengths, m):
        x = self.embedding(ids).transpose(1, 2)
        mm = m[:, None, :]
        x = x.masked_fill(mm, 0.0)
        for cnn in self.cnn:
            x = cnn(x).masked_fill(mm, 0.0)
        x = x.transpose(1, 2)
        packed = nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), batch_first=True, enforce_sorted=False)
        y, _ = self.lstm(packed)
        y, _ = nn.utils.rnn.pad_packed_sequence(y, batch_first=True)
        out = torch.zeros(y.shape[0], m.shape[-1], y.shape[2], device=y.device)
        out[:, : y.shape[1]] = y
        return out.transpose(1, 2).masked_fill(mm, 0.0)  # [B,512,T]


# --------------------------------------------------------------------------------------
# A.10-A.12 decoder / iSTFTNet generator
# --------------------------------------------------------------------------------------
def _pad(k, d):
    return (k * d - d) // 2


class AdaINResBlock1(nn.Module):
    """A.12"""

    def __init__(self, ch, k, dil, style_dim):
        super().__init__()
        self.convs1 = nn.ModuleList([weight_norm(nn.Conv1d(ch, ch, k, 1, dilation=d, padding=_pad(k, d))) for d in dil])
        self.convs2 = nn.ModuleList([weight_norm(nn.Conv1d(ch, ch, k, 1, dilation=1, padding=_pad(k, 1))) for _ in dil])
        self.adain1 = nn.ModuleList([AdaIN1d(style_dim, ch) for _ in dil])
        self.adain2 = nn.ModuleList([AdaIN1d(style_dim, ch) for _ in dil])
        self.alpha1 = nn.ParameterList([nn.Parameter(torch.ones(1, ch, 1)) for _ in dil])
        self.alpha2 = nn.ParameterList([nn.Parameter(torch.ones(1, ch, 1)) for _ in dil])

    def forward(self, x, s):
        for c1, c2, n1, n2, a1, a2 in zip(self.convs1, self.convs2, self.adain1, self.adain2, self.alpha1, self.alpha2):
            xt = n1(x, s)
            xt = xt + (1 / a1) * (torch.sin(a1 * xt) ** 2)
            xt = c1(xt)
            xt = n2(xt, s)
            xt = xt + (1 / a2) * (torch.sin(a2 * xt) ** 2)
            xt = c2(xt)
            x = xt + x
        return x


def harmonic_phase(f0: torch.Tensor, upsample_scale: int, sr: int, harmonics: int) -> torch.Tensor:
    """SineGen._f02sine up to (not including) sin(): f0 [B,L] -> phase [B, L*scale, H] (fp32).
    rand_ini omitted (no effect)."""
    f0u = F.interpolate(f0[:, None, :], scale_factor=float(upsample_scale), mode="nearest").transpose(1, 2)  # [B,N,1]
    fn = f0u * torch.arange(1, harmonics + 1, dtype=torch.float32, device=f0.device)[None, None, :]
    rad = (fn / sr) % 1
    rad = F.interpolate(rad.transpose(1, 2), scale_factor=1 / upsample_scale, mode="linear").transpose(1, 2)
    phase = torch.cumsum(rad, dim=1) * 2 * torch.pi
    phase = F.interpolate(phase.transpose(1, 2) * upsample_scale, scale_factor=float(upsample_scale), mode="linear").transpose(1, 2)
    return phase, f0u


class SourceModule(nn.Module):
    """SourceModuleHnNSF + SineGen (A.11 source), noise injected."""

    def __init__(self, sr, upsample_scale, harmonic_num=8, sine_amp=0.1, noise_std=0.003, voiced_threshold=10.0):
        super().__init__()
        self.sr, self.scale, self.H = sr, upsample_scale, harmonic_num + 1
        self.sine_amp, self.noise_std, self.thr = sine_amp, noise_std, voiced_threshold
        self.l_linear = nn.Linear(self.H, 1)

    def forward(self, f0, noise):  # f0 [B,L]; noise [B, L*scale, H] ~ N(0,1)
        phase, f0u = harmonic_phase(f0, self.scale, self.sr, self.H)
        sines = torch.sin(phase) * self.sine_amp
        uv = (f0u > self.thr).to(torch.float32)
        noise_amp = uv * self.noise_std + (1 - uv) * self.sine_amp / 3
        sine_waves = sines * uv + noise_amp * noise
        return torch.tanh(self.l_linear(sine_waves)).squeeze(-1)  # [B,N]


class Generator(nn.Module):
    def __init__(self, c: KokoroConfig):
        super().__init__()
        sd = c.style_dim
        self.c = c
        self.nk = len(c.resblock_kernel_sizes)
        self.nu = len(c.upsample_rates)
        scale = math.prod(c.upsample_rates) * c.gen_istft_hop_size
        self.m_source = SourceModule(c.sampling_rate, scale)  # key m_source.l_linear.{weight,bias}
        self.noise_convs = nn.ModuleList()
        self.noise_res = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch0 = c.upsample_initial_channel
        for i, (u, k) in enumerate(zip(c.upsample_rates, c.upsample_kernel_sizes)):
            self.ups.append(weight_norm(nn.ConvTranspose1d(ch0 // (2 ** i), ch0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)))
        for i in range(self.nu):
            ch = ch0 // (2 ** (i + 1))
            for k, d in zip(c.resblock_kernel_sizes, c.resblock_dilation_sizes):
                self.resblocks.append(AdaINResBlock1(ch, k, d, sd))
            if i + 1 < self.nu:
                sf = math.prod(c.upsample_rates[i + 1:])
                self.noise_convs.append(nn.Conv1d(c.gen_istft_n_fft + 2, ch, kernel_size=sf * 2, stride=sf, padding=(sf + 1) // 2))
                self.noise_res.append(AdaINResBlock1(ch, 7, (1, 3, 5), sd))
            else:
                self.noise_convs.append(nn.Conv1d(c.gen_istft_n_fft + 2, ch, kernel_size=1))
                self.noise_res.append(AdaINResBlock1(ch, 11, (1, 3, 5), sd))
        self.conv_post = weight_norm(nn.Conv1d(ch, c.gen_istft_n_fft + 2, 7, 1, padding=3))
        self.n_fft, self.hop = c.gen_istft_n_fft, c.gen_istft_hop_size
        self.register_buffer("window", torch.hann_window(self.n_fft, periodic=True), persistent=False)

    def stft(self, x):
        z = torch.stft(x, self.n_fft, self.hop, self.n_fft, window=self.window, return_complex=True)
        return torch.abs(z), torch.angle(z)

    def istft(self, mag, phase):
        return torch.istft(mag * torch.exp(phase * 1j), self.n_fft, self.hop, self.n_fft, window=self.window)

    def forward(self, x, s, f0, noise, taps=None, har_override=None):
        har_source = self.m_source(f0, noise)
        mag, ang = self.stft(har_source)
        har = torch.cat([mag, ang], dim=1)
        if taps is not None:
            taps["har_own"] = har
        if har_override is not None:
            # angle() is discontinuous: wherever Im X ~ 0 and Re X < 0 it is +pi or -pi by the SIGN of rounding noise (always at
            # STFT frame 0, which reflect padding makes exactly symmetric, and wherever a bin crosses the negative real axis).
            # That sign depends on the FFT library, not on the model.  Parity runs therefore (1) compare the implementation's
            # source spectrum with this one as COMPLEX numbers and (2) feed the implementation's [mag | angle] features here,
            # so the rest of the generator is compared on bit-identical inputs (tests/parity.py).
            har = har_override
        if taps is not None:
            taps["har_source"], taps["har"] = har_source, har
        for i in range(self.nu):
            x = F.leaky_relu(x, 0.1)
            xs = self.noise_res[i](self.noise_convs[i](har), s)
            x = self.ups[i](x)
            if i == self.nu - 1:
                x = F.pad(x, (1, 0), mode="reflect")
            x = x + xs
            acc = None
            for j in range(self.nk):
                y = self.resblocks[i * self.nk + j](x, s)
                acc = y if acc is None else acc + y
            x = acc / self.nk
            if taps is not None:
                taps[f"gen_stage{i}"] = x
        x = F.leaky_relu(x)
        x = self.conv_post(x)
        nb = self.n_fft // 2 + 1
        spec, phase = torch.exp(x[:, :nb]), torch.sin(x[:, nb:])
        if taps is not None:
            taps["conv_post"] = x
        return self.istft(spec, phase)[:, None, :]


class Decoder(nn.Module):
    """A.10 — called per sample at kmodel.py:143."""

    def __init__(self, c: KokoroConfig):
        super().__init__()
        d, sd = c.hidden_dim, c.style_dim
        self.encode = AdainResBlk1d(d + 2, 1024, sd)
        self.decode = nn.ModuleList([AdainResBlk1d(1024 + 2 + 64, 1024, sd) for _ in range(3)] + [AdainResBlk1d(1024 + 2 + 64, 512, sd, upsample=True)])
        self.F0_conv = weight_norm(nn.Conv1d(1, 1, 3, stride=2, padding=1))
        self.N_conv = weight_norm(nn.Conv1d(1, 1, 3, stride=2, padding=1))
        self.asr_res = nn.Sequential(weight_norm(nn.Conv1d(512, 64, 1)))
        self.generator = Generator(c)

    def forward(self, asr, F0_curve, N, s, noise, taps=None, har_override=None):
        F0 = self.F0_conv(F0_curve[:, None])
        Nn = self.N_conv(N[:, None])
        x = self.encode(torch.cat([asr, F0, Nn], dim=1), s)
        asr_res = self.asr_res(asr)
        res = True
        for blk in self.decode:
            if res:
                x = torch.cat([x, asr_res, F0, Nn], dim=1)
            x = blk(x, s)
            if blk.up:
                res = False
        if taps is not None:
            taps["gen_in"] = x
        return self.generator(x, s, F0_curve, noise, taps, har_override)


# --------------------------------------------------------------------------------------
# the model, wired as the reference wires it
# --------------------------------------------------------------------------------------
@dataclass
class OracleOutput:
    audio: List[torch.Tensor]
    pred_dur: List[torch.Tensor]
    indices: List[torch.Tensor]
    duration: List[torch.Tensor]          # pre-round fp32 value (margin-to-.5 diagnostics)
    taps: List[Dict[str, torch.Tensor]] = field(default_factory=list)


class KokoroOracle(nn.Module):
    def __init__(self, c: Optional[KokoroConfig] = None):
        super().__init__()
        c = c or KokoroConfig()
        self.c = c
        self.bert = Albert(c)
        self.bert_encoder = nn.Linear(c.bert_hidden, c.hidden_dim)
        self.predictor = ProsodyPredictor(c)
        self.text_encoder = TextEncoder(c)
        self.decoder = Decoder(c)

    # ---- checkpoint layout: {bert, bert_encoder, predictor, decoder, text_encoder} -> state dicts
    def export_checkpoint(self) -> Dict[str, Dict[str, torch.Tensor]]:
        out = {}
        for name in ("bert", "bert_encoder", "predictor", "decoder", "text_encoder"):
            sd = getattr(self, name).state_dict()
            # InstanceNorm affine weight/bias (=1/0) are not in the real checkpoint (loaded strict=False)
            out[name] = {"module." + k: v.detach().clone() for k, v in sd.items() if ".norm.weight" not in k and ".norm.bias" not in k}
        return out

    def load_checkpoint(self, ckpt: Dict[str, Dict[str, torch.Tensor]]):
        for name, sd in ckpt.items():
            sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
            getattr(self, name).load_state_dict(sd, strict=False)
        return self

    def noise_shape(self, F_frames: int):
        return (600 * F_frames, 9)

    @torch.no_grad()
    def front(self, padded_ids, input_lengths, ref_s, speeds):
        """Batched front half, kmodel.py:62-103 (incl. the un-packed predictor.lstm, App. C-1)."""
        B, Tm = padded_ids.shape
        text_mask = torch.arange(Tm, device=padded_ids.device)[None].expand(B, -1) + 1 > input_lengths[:, None]
        bert_dur = self.bert(padded_ids, attention_mask=(~text_mask).int())
        d_en = self.bert_encoder(bert_dur).transpose(-1, -2)
        style = ref_s[:, 128:]
        d = self.predictor.text_encoder(d_en, style, input_lengths, text_mask)
        x, _ = self.predictor.lstm(d)
        duration = torch.sigmoid(self.predictor.duration_proj(x)).sum(dim=-1)
        duration = duration.clone()
        for i, sp in enumerate(speeds):
            duration[i] = duration[i] / sp
        return d, duration, text_mask

    @torch.no_grad()
    def forward_batch(self, ids_list: List[List[int]], ref_s: torch.Tensor, speeds: Optional[List[float]] = None,
                      noise: Optional[List[torch.Tensor]] = None, noise_seed: int = 0, want_taps: bool = False,
                      stop_after_durations: bool = False, tap_keys=None) -> OracleOutput:
        """``ids_list[b]`` already includes BOS/EOS zeros (kmodel.py:36-42).  ``tap_keys``: keep only these taps (a batch-32
        run with every generator stage kept is ~1 GB)."""
        B = len(ids_list)
        speeds = speeds or [1.0] * B
        if ref_s.dim() == 3:
            ref_s = ref_s.squeeze(1)
        dev = self.bert_encoder.weight.device       # cpu, or a GPU for bench.py's eager-PyTorch baseline (same graph, stock kernels)
        ref_s = ref_s.to(dev)
        lengths = torch.tensor([len(x) for x in ids_list], dtype=torch.long, device=dev)
        Tm = int(lengths.max())
        padded = torch.zeros(B, Tm, dtype=torch.long, device=dev)
        for i, ids in enumerate(ids_list):
            padded[i, : len(ids)] = torch.tensor(ids, dtype=torch.long, device=dev)
        d, duration, text_mask = self.front(padded, lengths, ref_s, speeds)
        out = OracleOutput([], [], [], [])
        for i in range(B):  # kmodel.py:108-144
            L = int(lengths[i])
            cur = duration[i, :L]
            pred_dur = torch.round(cur).clamp(min=1).long()
            indices = torch.repeat_interleave(torch.arange(L, device=dev), pred_dur)
            out.duration.append(cur.clone())
            out.pred_dur.append(pred_dur)
            out.indices.append(indices)
            if stop_after_durations:
                continue
            Fr = int(indices.shape[0])
            aln = torch.zeros(L, Fr, device=dev)
            aln[indices, torch.arange(Fr, device=dev)] = 1
            aln = aln[None]
            s_i = ref_s[i: i + 1, 128:]
            en = d[i: i + 1, :L].transpose(-1, -2) @ aln
            F0, N = self.predictor.F0Ntrain(en, s_i)
            t_en = self.text_encoder(padded[i: i + 1, :L], torch.tensor([L]), text_mask[i: i + 1, :L])     # lengths stay on the CPU (packing)
            asr = t_en @ aln
            if noise is not None:
                nz = noise[i].to(dev)
            else:
                nz = make_noise(noise_seed, i, Fr).to(dev)
            taps = {} if want_taps else None
            audio = self.decoder(asr, F0, N, ref_s[i: i + 1, :128], nz[None], taps).squeeze()
            if want_taps:
                taps.update(dict(d=d[i, :L], en=en[0], F0=F0[0], N=N[0], t_en=t_en[0], asr=asr[0]))
                if tap_keys is not None:
                    taps = {k: v for k, v in taps.items() if k in tap_keys}
                out.taps.append(taps)
            out.audio.append(audio)
        return out

    @torch.no_grad()
    def vocoder(self, x, s, f0, noise, taps=None, har_override=None):
        """Generator only (BASELINE config 4): x [B,512,L], s [B,128], f0 [B,L], noise [B,300L,9]."""
        return self.decoder.generator(x, s, f0, noise, taps, har_override).squeeze(1)


def make_noise(seed: int, utt: int, frames: int) -> torch.Tensor:
    """The injected SineGen noise for utterance ``utt``: N(0,1) [600*frames, 9], reproducible."""
    g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + utt)
    return torch.randn(600 * frames, 9, generator=g)


# --------------------------------------------------------------------------------------
# host-side helpers mirroring kmodel.py:36-42 and pipeline.py:199,596-601
# --------------------------------------------------------------------------------------
def phonemes_to_ids(vocab: Dict[str, int], phonemes: str) -> List[int]:
    ids = [vocab.get(p) for p in phonemes]
    return [0] + [i for i in ids if i is not None] + [0]
