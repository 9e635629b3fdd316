"""Drop-in mirror of the reference's engine class.

Reference interface (unchanged callers): ``BatchKModel`` in /root/reference/src/illufly_tts/core/kmodel.py:5-150,
constructed at core/pipeline.py:70 as ``BatchKModel(repo_id=repo_id).to(device).eval()``, used at
core/pipeline.py:204 (``model(phonemes, ref_s, speed, return_output=True).audio``) and core/pipeline.py:607
(``model.forward_batch(ipa_list, ref_s_batch, speeds)``); ``.vocab`` and ``.device`` are read at kmodel.py:38,58.

Same names, argument meaning and error behaviour; the arithmetic runs in libttsb200.so on one B200.
"""
from __future__ import annotations

import json
import logging
import os
import sys
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence, Union

import torch

from .engine import Engine, PRECISION_FP32

logger = logging.getLogger(__name__)
SILENCE_SAMPLES = 24000          # kmodel.py:30,91,150
MAX_TOKENS = 512                 # ALBERT positions; kokoro KModel.forward asserts len(ids)+2 <= 512


def phonemes_to_ids(vocab: Mapping[str, int], phonemes: str) -> List[int]:
    """kmodel.py:36-42 — map characters through the vocab, drop unknown ones, wrap in BOS/EOS = 0."""
    ids = [vocab.get(p) for p in phonemes]
    return [0] + [i for i in ids if i is not None] + [0]


@dataclass
class Output:
    """kokoro ``KModel.Output``: what ``__call__(..., return_output=True)`` returns (pipeline.py:204-206 reads .audio)."""
    audio: torch.Tensor
    pred_dur: Optional[torch.Tensor] = None


def _load_checkpoint(model: Union[str, Mapping, None], config: Union[str, Mapping, None], repo_id: Optional[str]):
    """Resolve (config dict, checkpoint dict) the way kokoro's KModel.__init__ does, minus the network:
    explicit arguments first, then $ILLUFLY_TTS_B200_MODEL_DIR / a local directory named by repo_id, then the HF cache."""
    search = []
    if isinstance(model, str) and os.path.isdir(model):
        search.append(model)
        model = None
    if os.environ.get("ILLUFLY_TTS_B200_MODEL_DIR"):
        search.append(os.environ["ILLUFLY_TTS_B200_MODEL_DIR"])
    if repo_id and os.path.isdir(repo_id):
        search.append(repo_id)
    if config is None or model is None:
        for d in search:
            if config is None and os.path.exists(os.path.join(d, "config.json")):
                config = os.path.join(d, "config.json")
            if model is None:
                pth = sorted(f for f in os.listdir(d) if f.endswith(".pth"))
                if pth:
                    model = os.path.join(d, pth[0])
    if (config is None or model is None) and repo_id:
        try:  # same files kokoro fetches (config.json + kokoro-v1_1-zh.pth); only works with a populated HF cache offline
            from huggingface_hub import hf_hub_download
            if config is None:
                config = hf_hub_download(repo_id=repo_id, filename="config.json")
            if model is None:
                name = {"hexgrad/Kokoro-82M": "kokoro-v1_0.pth", "hexgrad/Kokoro-82M-v1.1-zh": "kokoro-v1_1-zh.pth"}.get(repo_id, "kokoro-v1_1-zh.pth")
                model = hf_hub_download(repo_id=repo_id, filename=name)
        except Exception as exc:  # no network / not cached
            raise FileNotFoundError(
                f"cannot locate config.json / checkpoint for repo_id={repo_id!r}: pass config=/model=, or set "
                f"ILLUFLY_TTS_B200_MODEL_DIR to a directory holding config.json and the .pth ({exc})") from exc
    if config is None or model is None:
        raise FileNotFoundError("BatchKModel needs a config (dict or json path) and a checkpoint (dict or .pth path)")
    if isinstance(config, str):
        with open(config, "r", encoding="utf-8") as f:
            config = json.load(f)
    if isinstance(model, str):
        model = torch.load(model, map_location="cpu", weights_only=True)
    return dict(config), model


class BatchKModel:
    """Same surface as the reference's ``BatchKModel(KModel)``.

    ``strict=False`` (default) keeps the reference's error surface: ``forward_batch`` never raises, failures give
    ``zeros(24000)`` per item and a too-short phoneme string gives a ONE-element list (kmodel.py:27-30, 147-150).
    ``strict=True`` raises instead (tests).  Extra keyword-only arguments (``noise``, ``seed``) inject the SineGen noise
    for parity runs; the reference draws it from the unseeded global RNG.

    Deviation, on purpose: for a ragged batch the reference's ``curr_d.transpose(-1,-2) @ pred_aln_trg`` (kmodel.py:132)
    multiplies a [640,Tmax] by a [T_b,F] matrix, raises, and the whole batch becomes silence.  This class synthesises the
    utterances instead (``d_batch[i,:T_b]``, what the code means); ``ragged_silence=True`` restores the reference behaviour.
    """

    def __init__(self, repo_id: Optional[str] = None, config: Union[Dict, str, None] = None, model: Union[str, Mapping, None] = None,
                 disable_complex: bool = False, *, precision: int = PRECISION_FP32, strict: bool = False, ragged_silence: bool = False,
                 padded_lstm: bool = True):
        self.repo_id = repo_id
        cfg, ckpt = _load_checkpoint(model, config, repo_id)
        self.vocab: Dict[str, int] = cfg["vocab"]
        self.context_length = int(cfg.get("plbert", {}).get("max_position_embeddings", MAX_TOKENS))
        self._ckpt = ckpt
        self._engine: Optional[Engine] = None
        self._device = torch.device("cpu")
        self.precision = precision
        self.strict = strict
        self.ragged_silence = ragged_silence
        self.padded_lstm = padded_lstm
        self.training = False

    # ---- nn.Module-like surface the pipeline touches (pipeline.py:70) -----------------------------------
    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"illufly-tts_b200 BatchKModel runs on CUDA (B200) only; got device {dev}. There is no CPU fallback.")
        if self._engine is not None and self._device == dev:
            return self
        if self._engine is not None:
            self._engine.close()
        self._engine = Engine(self._ckpt, dev, self.precision)
        self.__dict__.pop("_voice_slots", None)          # voice packs live in the engine: a new engine starts with none
        self._device = self._engine.device
        return self

    def cuda(self, device=0):
        return self.to(f"cuda:{device}" if isinstance(device, int) else device)

    def eval(self):
        self.training = False
        return self

    @property
    def device(self):
        return self._device

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise RuntimeError("BatchKModel has no engine yet: call .to('cuda:N') first (the reference does so at pipeline.py:70)")
        return self._engine

    # ---- kokoro KModel.forward / __call__ (pipeline.py:204) ---------------------------------------------
    @torch.no_grad()
    def __call__(self, phonemes: str, ref_s: torch.Tensor, speed: float = 1, return_output: bool = False, *, noise=None, seed: int = 0):
        input_ids = phonemes_to_ids(self.vocab, phonemes)
        assert len(input_ids) <= self.context_length, (len(input_ids), self.context_length)     # kokoro: len(input_ids)+2 <= context_length
        ref = ref_s.reshape(1, -1)
        res = self.engine.synthesize([input_ids], ref, [float(speed)], noise=None if noise is None else [noise], seed=seed, padded_lstm=False)
        audio = res.audio[0].cpu()               # KModel.forward: audio.squeeze().cpu()
        pred_dur = res.pred_dur[0]
        return Output(audio=audio, pred_dur=pred_dur) if return_output else audio

    forward = __call__

    # ---- BatchKModel.forward_batch (kmodel.py:7-150) ------------------------------------------------------
    @torch.no_grad()
    def forward_batch(self, phonemes_batch: Sequence[str], ref_s_batch: torch.Tensor, speeds: Optional[Sequence[float]] = None, *,
                      noise=None, seed: int = 0):
        if speeds is None:
            speeds = [1.0] * len(phonemes_batch)
        for i, ph in enumerate(phonemes_batch):                           # kmodel.py:27-30
            if not ph or len(ph) < 5:
                logger.error("phoneme sequence #%d is empty or too short: %r", i, ph)
                if self.strict:
                    raise ValueError(f"phoneme sequence #{i} is empty or too short")
                return [torch.zeros(SILENCE_SAMPLES)]
        B = len(phonemes_batch)
        try:
            ids_list = [phonemes_to_ids(self.vocab, ph) for ph in phonemes_batch]
            if self.ragged_silence and len({len(x) for x in ids_list}) > 1:
                raise RuntimeError("ragged batch (reference kmodel.py:132 raises on it)")
            res = self.engine.synthesize(ids_list, ref_s_batch, [float(s) for s in speeds], noise=noise, seed=seed, padded_lstm=self.padded_lstm)
            return res.audio                                              # list of 1-D fp32 tensors on self.device
        except Exception as exc:                                          # kmodel.py:147-150
            if self.strict:
                raise
            logger.error("batched synthesis failed: %s", exc, exc_info=True)
            return [torch.zeros(SILENCE_SAMPLES) for _ in range(B)]


    # ---- voice packs resident on the device (the reference: pipeline.py:78-109 cache, :596-603 row select) ----
    def register_voice(self, voice_id: str, pack: torch.Tensor) -> int:
        """Upload ``pack`` ([510,1,256]: ``pipeline.voices[voice_id]``) to the engine once; returns its slot.  Registering the same id
        again with the same tensor object is a no-op, with another tensor it replaces the pack."""
        reg = self.__dict__.setdefault("_voice_slots", {})
        ent = reg.get(voice_id)
        if ent is not None and ent[1] is pack:
            return ent[0]
        slot = ent[0] if ent is not None else len(reg)
        rows = self.engine.set_voice_pack(slot, pack)
        reg[voice_id] = (slot, pack, rows)
        return slot

    @torch.no_grad()
    def forward_batch_voices(self, phonemes_batch: Sequence[str], voice_ids: Sequence[str], speeds: Optional[Sequence[float]] = None, *,
                             noise=None, seed: int = 0):
        """``forward_batch`` with the style rows taken from registered voice packs on the device: utterance b uses row
        ``len(phonemes_batch[b]) - 1`` of ``voice_ids[b]``'s pack (pipeline.py:596-601), clamped to the pack.  Same results, bit for bit,
        as ``forward_batch`` with ``torch.stack([pack[len(ps) - 1] ...])``; no style bytes cross PCIe per batch.  Error surface of
        ``forward_batch`` (kmodel.py:27-30, 147-150)."""
        if speeds is None:
            speeds = [1.0] * len(phonemes_batch)
        for i, ph in enumerate(phonemes_batch):
            if not ph or len(ph) < 5:
                logger.error("phoneme sequence #%d is empty or too short: %r", i, ph)
                if self.strict:
                    raise ValueError(f"phoneme sequence #{i} is empty or too short")
                return [torch.zeros(SILENCE_SAMPLES)]
        B = len(phonemes_batch)
        try:
            reg = self.__dict__.get("_voice_slots", {})
            slots, rows = [], []
            for ph, vid in zip(phonemes_batch, voice_ids):
                if vid not in reg:
                    raise KeyError(f"voice {vid!r} is not registered (register_voice)")
                slot, _, nrows = reg[vid]
                slots.append(slot)
                rows.append(min(max(len(ph), 1) - 1, nrows - 1))
            ids_list = [phonemes_to_ids(self.vocab, ph) for ph in phonemes_batch]
            if self.ragged_silence and len({len(x) for x in ids_list}) > 1:
                raise RuntimeError("ragged batch (reference kmodel.py:132 raises on it)")
            res = self.engine.synthesize(ids_list, None, [float(s) for s in speeds], noise=noise, seed=seed, padded_lstm=self.padded_lstm,
                                         voice_slots=slots, voice_rows=rows)
            return res.audio
        except Exception as exc:
            if self.strict:
                raise
            logger.error("batched synthesis failed: %s", exc, exc_info=True)
            return [torch.zeros(SILENCE_SAMPLES) for _ in range(B)]


def install_into_reference():
    """Make ``illufly_tts.core.kmodel.BatchKModel`` resolve to this class, so the unmodified reference pipeline
    (core/pipeline.py:19 ``from .kmodel import BatchKModel``) constructs the B200 engine.  The reference tree is not
    touched: a stand-in module is registered in ``sys.modules`` before the reference is imported."""
    import types
    mod = types.ModuleType("illufly_tts.core.kmodel")
    mod.BatchKModel = BatchKModel
    mod.__doc__ = "illufly-tts_b200 stand-in for illufly_tts.core.kmodel"
    sys.modules["illufly_tts.core.kmodel"] = mod
    return mod
