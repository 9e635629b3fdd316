"""Host wrapper over the C ABI (include/ttsb200.h).

PyTorch is used here only for device memory (output tensors), checkpoint I/O and the caller-facing tensor types;
all arithmetic happens in libttsb200.so.  Nothing in this module computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import functools
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib

PRECISION_FP32 = 0          # the strict tier: tcgen05 with split-fp16 operands (fp32-class products) on every convolution
PRECISION_BF16 = 1
PRECISION_FP16 = 2
PRECISION_FP32_FFMA = 3     # CUDA-core fp32 everywhere: the arithmetic reference of the strict tier
CHECKPOINT_MODULES = ("bert", "bert_encoder", "predictor", "decoder", "text_encoder")


@dataclass
class SynthesisResult:
    audio: List[torch.Tensor]                 # per utterance, 1-D float32 on the engine device (or CPU if asked)
    pred_dur: List[torch.Tensor]              # per utterance int64 [T] (CPU)
    duration: List[torch.Tensor]              # per utterance fp32 [T] pre-round value (CPU)
    indices: List[torch.Tensor]               # per utterance int64 [F] alignment index map (CPU); empty unless asked
    frame_offsets: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, torch.Tensor):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _serialised(method):
    """One engine = one CUDA stream + one workspace, and a synthesis is a begin -> finish PAIR with state in between: callers
    on different threads (the reference reaches its model through ``asyncio.to_thread`` from several places) take turns.
    Without this, a second thread's begin() would reset the workspace under the first thread's finish(): wrong audio, silently."""
    @functools.wraps(method)
    def locked(self, *a, **kw):
        with self._lock:
            return method(self, *a, **kw)
    return locked


class Engine:
    """One engine = one GPU.  Mirrors what kokoro's ``KModel.__init__`` + ``.to(device)`` do for the reference:
    take the checkpoint ``{bert, bert_encoder, predictor, decoder, text_encoder}`` of state dicts (``module.``
    prefixes allowed, old-style ``weight_g``/``weight_v``), hand every tensor to the library, finalize."""

    def __init__(self, checkpoint: Mapping[str, Mapping[str, torch.Tensor]], device: int | str | torch.device = 0,
                 precision: int = PRECISION_FP32):
        self._lock = threading.RLock()
        self._lib = _lib.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError(f"illufly-tts_b200 runs on CUDA devices only (got {dev}); there is no CPU fallback")
        self.device = torch.device("cuda", dev.index or 0)
        h = C.c_void_p()
        _lib.check(self._lib.ttsb_engine_create(self.device.index, C.byref(h)))
        self._h = h
        try:
            for mod in CHECKPOINT_MODULES:
                if mod not in checkpoint:
                    raise KeyError(f"checkpoint has no '{mod}' state dict")
                for key, t in checkpoint[mod].items():
                    k = key[7:] if key.startswith("module.") else key
                    if not torch.is_tensor(t) or not t.dtype.is_floating_point:
                        continue                       # e.g. position_ids buffers of older transformers checkpoints
                    a = np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy())
                    if a.ndim == 0:
                        a = a.reshape(1)
                    shape = (C.c_int64 * a.ndim)(*a.shape)
                    _lib.check(self._lib.ttsb_engine_set_tensor(self._h, f"{mod}.{k}".encode(), _ptr(a), shape, a.ndim))
            _lib.check(self._lib.ttsb_engine_finalize(self._h))
            self.set_precision(precision)
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttsb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @_serialised
    def set_precision(self, precision: int):
        _lib.check(self._lib.ttsb_engine_set_precision(self._h, int(precision)))

    @property
    def precision(self) -> int:
        return int(self._lib.ttsb_engine_get_precision(self._h))

    @_serialised
    def enable_taps(self, on: bool = True):
        _lib.check(self._lib.ttsb_engine_enable_taps(self._h, int(on)))

    def launch_count(self) -> int:
        return int(self._lib.ttsb_engine_launch_count(self._h))

    def last_step_ms(self) -> float:
        ms = C.c_float(0)
        _lib.check(self._lib.ttsb_engine_last_step_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def workspace_peak_bytes(self) -> int:
        """high-water mark of the engine's workspace during the last step"""
        n = C.c_int64(0)
        _lib.check(self._lib.ttsb_engine_workspace_peak(self._h, C.byref(n)))
        return int(n.value)

    PROFILE_CATEGORIES = ("conv_tensor", "conv_fp32", "lstm", "istft", "upsample_add", "gather", "attention", "layernorm", "harmonic_source",
                          "fp32_projections")

    @_serialised
    def profile(self, on: bool = True):
        _lib.check(self._lib.ttsb_engine_profile(self._h, int(on)))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for i, name in enumerate(self.PROFILE_CATEGORIES):
            ms, fl, by, n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
            _lib.check(self._lib.ttsb_engine_profile_read(self._h, i, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)))
            out[name] = dict(ms=ms.value, flops=fl.value, bytes=by.value, launches=int(n.value))
        return out

    @_serialised
    def read_tap(self, name: str) -> torch.Tensor:
        rows, cols = C.c_int64(), C.c_int64()
        _lib.check(self._lib.ttsb_engine_read_tap(self._h, name.encode(), None, C.byref(rows), C.byref(cols)))
        out = np.empty((rows.value, cols.value), np.float32)
        _lib.check(self._lib.ttsb_engine_read_tap(self._h, name.encode(), _ptr(out), C.byref(rows), C.byref(cols)))
        return torch.from_numpy(out)

    # ------------------------------------------------------------------------------------------------
    @_serialised
    def set_voice_pack(self, slot: int, pack: Optional[torch.Tensor]) -> int:
        """Upload a voice pack ([rows,1,256] or [rows,256], the reference's ``voices[voice_id]``, pipeline.py:78-109) into device slot
        ``slot`` - once; ``synthesize(voice_slots=..., voice_rows=...)`` then picks style rows on the device.
        ``pack=None`` empties the slot.  Returns the number of rows."""
        if pack is None:
            _lib.check(self._lib.ttsb_engine_voice_pack_set(self._h, int(slot), None, 0))
            return 0
        p = pack.detach().to(torch.float32).reshape(pack.shape[0], -1).cpu()
        if p.shape[1] != 256:
            raise ValueError(f"a voice pack is [rows,1,256] or [rows,256]; got {tuple(pack.shape)}")
        arr = np.ascontiguousarray(p.numpy())
        _lib.check(self._lib.ttsb_engine_voice_pack_set(self._h, int(slot), _ptr(arr), int(arr.shape[0])))
        return int(arr.shape[0])

    @_serialised
    def synthesize(self, ids_list: Sequence[Sequence[int]], ref_s: Optional[torch.Tensor], speeds: Optional[Sequence[float]] = None,
                   noise: Optional[Sequence[torch.Tensor]] = None, seed: int = 0, padded_lstm: bool = True,
                   want_indices: bool = False, forced_dur: Optional[Sequence[Sequence[int]]] = None,
                   audio_to_host: bool = False, voice_slots: Optional[Sequence[int]] = None,
                   voice_rows: Optional[Sequence[int]] = None) -> SynthesisResult:
        """ids_list[b] already carries BOS/EOS (kmodel.py:36-42).  ref_s [B,256] or [B,1,256] (any device) - or, with resident voice
        packs (``set_voice_pack``), ``voice_slots`` / ``voice_rows`` per utterance and ``ref_s=None``."""
        B = len(ids_list)
        if B == 0:
            return SynthesisResult([], [], [], [])
        lens = np.array([len(x) for x in ids_list], np.int32)
        tok_off = np.zeros(B + 1, np.int32)
        np.cumsum(lens, out=tok_off[1:])
        ids = np.concatenate([np.asarray(x, np.int32) for x in ids_list]).astype(np.int32, copy=False)
        by_voice = voice_slots is not None
        if by_voice:
            if voice_rows is None or len(voice_slots) != B or len(voice_rows) != B:
                raise ValueError("voice_slots and voice_rows must both give one entry per utterance")
            vs = np.asarray(list(voice_slots), np.int32)
            vr = np.asarray(list(voice_rows), np.int32)
        else:
            if ref_s is None:
                raise ValueError("ref_s is required unless voice_slots / voice_rows are given")
            ref = ref_s.detach()
            if ref.dim() == 3:
                ref = ref.squeeze(1)
            ref = np.ascontiguousarray(ref.to(torch.float32).cpu().numpy())
            if ref.shape != (B, 256):
                raise ValueError(f"ref_s must be [B,256] or [B,1,256]; got {tuple(ref_s.shape)} for B={B}")
        sp = np.asarray([1.0] * B if speeds is None else list(speeds), np.float32)
        R = int(tok_off[-1])
        pred = np.empty(R, np.int32)
        dur = np.empty(R, np.float32)
        f_off = np.zeros(B + 1, np.int32)
        forced = None
        if forced_dur is not None:
            forced = np.concatenate([np.asarray(x, np.int32) for x in forced_dur]).astype(np.int32, copy=False)
            if forced.shape[0] != R:
                raise ValueError("forced_dur must give one duration per token")
        if by_voice:
            _lib.check(self._lib.ttsb_synthesize_begin_voices(self._h, _ptr(ids), _ptr(tok_off), B, _ptr(vs), _ptr(vr), _ptr(sp), int(padded_lstm),
                                                              _ptr(forced), _ptr(pred), _ptr(dur), _ptr(f_off)))
        else:
            _lib.check(self._lib.ttsb_synthesize_begin(self._h, _ptr(ids), _ptr(tok_off), B, _ptr(ref), _ptr(sp), int(padded_lstm),
                                                       _ptr(forced), _ptr(pred), _ptr(dur), _ptr(f_off)))
        RF = int(f_off[-1])
        nz, nz_mem = None, _lib.MEM_HOST
        if noise is not None:
            nz = noise if torch.is_tensor(noise) else torch.cat([n.reshape(-1, 9) for n in noise], 0)
            nz = nz.to(torch.float32).contiguous()
            if nz.numel() != RF * 600 * 9:
                raise ValueError(f"noise must hold 600*F*9 values per utterance ({RF * 5400} total), got {nz.numel()}")
            nz_mem = _lib.MEM_DEVICE if nz.is_cuda else _lib.MEM_HOST
            if nz.is_cuda:
                torch.cuda.synchronize(nz.device)      # produced on torch's stream, consumed on the engine's
        idx = np.empty(RF, np.int32) if want_indices else None
        if audio_to_host:
            # engine-owned pinned staging buffer, reused across calls: the returned views are valid until the next call
            if getattr(self, "_pinned_audio", None) is None or self._pinned_audio.numel() < RF * 600:
                self._pinned_audio = torch.empty(max(RF * 600, 1 << 20), dtype=torch.float32, pin_memory=True)
            audio = self._pinned_audio[: RF * 600]
            mem = _lib.MEM_HOST
        else:
            audio = torch.empty(RF * 600, dtype=torch.float32, device=self.device)
            mem = _lib.MEM_DEVICE
        _lib.check(self._lib.ttsb_synthesize_finish(self._h, _ptr(nz), nz_mem, C.c_uint64(seed & (2 ** 64 - 1)), _ptr(idx), _ptr(audio), mem))
        if not audio_to_host:
            # the engine runs on its own stream: make the result safe to use from torch's current stream
            torch.cuda.synchronize(self.device)
        res = SynthesisResult([], [], [], [], f_off)
        for b in range(B):
            res.audio.append(audio[600 * int(f_off[b]): 600 * int(f_off[b + 1])])
            res.pred_dur.append(torch.from_numpy(pred[tok_off[b]: tok_off[b + 1]].astype(np.int64)))
            res.duration.append(torch.from_numpy(dur[tok_off[b]: tok_off[b + 1]].copy()))
            if want_indices:
                res.indices.append(torch.from_numpy(idx[f_off[b]: f_off[b + 1]].astype(np.int64)))
        return res

    @_serialised
    def vocoder(self, x: torch.Tensor, s: torch.Tensor, f0: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
                lengths: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
        """Generator only (BASELINE config 4), kokoro layout in: x [B,512,L], s [B,128], f0 [B,L], noise [B,300L,9].
        Returns per-utterance audio [300 L_b] on the engine device."""
        B, _, L = x.shape
        lens = [L] * B if lengths is None else list(lengths)
        off = np.zeros(B + 1, np.int32)
        np.cumsum(np.asarray(lens, np.int32), out=off[1:])
        dev = self.device
        xr = torch.cat([x[b, :, : lens[b]].t() for b in range(B)], 0).to(dev, torch.float32).contiguous()      # [sum L, 512]
        fr = torch.cat([f0[b, : lens[b]] for b in range(B)], 0).to(dev, torch.float32).contiguous()
        sr = s.to(dev, torch.float32).contiguous()
        nz = None
        if noise is not None:
            nz = torch.cat([noise[b, : 300 * lens[b]] for b in range(B)], 0).to(dev, torch.float32).contiguous()
        RL = int(off[-1])
        audio = torch.empty(RL * 300, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        _lib.check(self._lib.ttsb_vocoder(self._h, _ptr(xr), _ptr(fr), _ptr(sr), _ptr(off), B, _lib.MEM_DEVICE, _ptr(nz),
                                          C.c_uint64(seed), _ptr(audio), _lib.MEM_DEVICE))
        torch.cuda.synchronize(dev)
        return [audio[300 * int(off[b]): 300 * int(off[b + 1])] for b in range(B)]

    # ---- single kernels (unit parity / roofline) ------------------------------------------------------
    def _dev_arg(self, name: str, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        """The op_* entry points hand raw pointers to kernels: a CPU tensor would be a wild device pointer (illegal address, and the
        CUDA context is gone), a float64 or strided one would be read as garbage.  Refuse them here with a message instead."""
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != self.device:
            where = getattr(t, "device", type(t).__name__)
            raise ValueError(f"{name} must be a tensor on {self.device} (got {where}): move it with .to('{self.device}')")
        if t.dtype != dtype:
            raise ValueError(f"{name} must be {dtype} (got {t.dtype})")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        return t

    @_serialised
    def op_istft(self, cp: torch.Tensor, frames_F: Sequence[int], timed: bool = False):
        """cp [sum(120F+1), ld>=22] device fp32 -> audio [600 sum F]"""
        self._dev_arg("cp", cp, torch.float32)
        fc = np.asarray(frames_F, np.int32)
        audio = torch.empty(600 * int(fc.sum()), dtype=torch.float32, device=self.device)
        ms = C.c_float(0)
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.ttsb_op_istft(self._h, _ptr(cp), cp.shape[1], _ptr(fc), len(fc), _ptr(audio), C.byref(ms)))
        return (audio, ms.value) if timed else audio

    @_serialised
    def op_adain_snake(self, x: torch.Tensor, row_offsets: Sequence[int], gamma_beta: torch.Tensor, alpha: torch.Tensor, timed: bool = False):
        self._dev_arg("x", x, torch.float32); self._dev_arg("gamma_beta", gamma_beta, torch.float32); self._dev_arg("alpha", alpha, torch.float32)
        off = np.asarray(row_offsets, np.int32)
        y = torch.empty_like(x)
        ms0, ms1 = C.c_float(0), C.c_float(0)
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.ttsb_op_adain_snake(self._h, _ptr(x), _ptr(y), x.shape[1], _ptr(off), len(off) - 1, _ptr(gamma_beta), _ptr(alpha),
                                                 C.byref(ms0), C.byref(ms1)))
        return (y, ms0.value, ms1.value) if timed else y

    @_serialised
    def op_length_regulate(self, pred_dur: torch.Tensor, token_offsets: Sequence[int], src: torch.Tensor, total_frames: int, timed: bool = False):
        self._dev_arg("pred_dur", pred_dur, torch.int32); self._dev_arg("src", src, torch.float32)
        off = np.asarray(token_offsets, np.int32)
        B = len(off) - 1
        idx = torch.empty(total_frames, dtype=torch.int32, device=self.device)
        dst = torch.empty(total_frames, src.shape[1], dtype=torch.float32, device=self.device)
        f_off = np.zeros(B + 1, np.int32)
        ms = C.c_float(0)
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.ttsb_op_length_regulate(self._h, _ptr(pred_dur), _ptr(off), B, _ptr(src), src.shape[1], _ptr(idx), _ptr(dst), _ptr(f_off),
                                                     C.byref(ms)))
        return (idx, dst, f_off, ms.value) if timed else (idx, dst, f_off)

    @_serialised
    def op_conv1d(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], row_offsets: Sequence[int], dilation: int = 1,
                  precision: int = PRECISION_FP32, reps: int = 1, timed: bool = False):
        """x [sum rows, Cin] device; w [Cout,Cin,k] (torch layout, any device); 'same' zero padding per utterance."""
        self._dev_arg("x", x, torch.float32)
        off = np.asarray(row_offsets, np.int32)
        wn = np.ascontiguousarray(w.detach().to(torch.float32).cpu().numpy())
        bn = None if bias is None else np.ascontiguousarray(bias.detach().to(torch.float32).cpu().numpy())
        Cout, Cin, k = wn.shape
        y = torch.empty(x.shape[0], Cout, dtype=torch.float32, device=self.device)
        ms = C.c_float(0)
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.ttsb_op_conv1d(self._h, _ptr(x), _ptr(wn), _ptr(bn), _ptr(y), Cin, Cout, k, dilation, _ptr(off), len(off) - 1,
                                            int(precision), reps, C.byref(ms)))
        return (y, ms.value) if timed else y

    @_serialised
    def op_bilstm(self, x: torch.Tensor, lstm_state: Mapping[str, torch.Tensor], row_offsets: Sequence[int], timed: bool = False):
        """x [sum rows, In] device; lstm_state = state_dict of an nn.LSTM(In, 256, bidirectional=True)."""
        self._dev_arg("x", x, torch.float32)
        off = np.asarray(row_offsets, np.int32)

        def cat(name):
            return np.ascontiguousarray(torch.cat([lstm_state[name + "_l0"], lstm_state[name + "_l0_reverse"]], 0).detach().to(torch.float32).cpu().numpy())
        wih, whh, bih, bhh = cat("weight_ih"), cat("weight_hh"), cat("bias_ih"), cat("bias_hh")
        y = torch.empty(x.shape[0], 512, dtype=torch.float32, device=self.device)
        ms = C.c_float(0)
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.ttsb_op_bilstm(self._h, _ptr(x), x.shape[1], _ptr(wih), _ptr(whh), _ptr(bih), _ptr(bhh), _ptr(off), len(off) - 1, _ptr(y),
                                            C.byref(ms)))
        return (y, ms.value) if timed else y
