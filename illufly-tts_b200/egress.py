"""Audio egress: waveform -> the bytes the reference's HTTP layer returns, without the disk round trip.

The reference writes every finished task with ``torchaudio.save(path, audio.float()[None], 24000)``
(/root/reference/src/illufly_tts/core/service.py:373-404), and the endpoint then re-reads the file and base64-encodes it
(api/endpoints.py:125-160).  What that call emits is a 32-bit IEEE-float mono RIFF/WAVE file.  The only byte-level evidence in
the reference tree is the response logged at service.log:33: a 650,480-byte file (RIFF size field 0x0009ECE8, base64 prefix
``UklGRujsCQ``).  Audio lengths of this engine are multiples of 600 samples; of the float-WAV header layouts in use (44 bytes:
bare ``fmt `` + ``data``; 58: ``fmt `` with cbSize + ``fact``; 80: ``fmt `` + ``fact`` + ``PEAK``, what libsndfile writes for
float data), only the 80-byte one gives a whole number of 600-sample frames (162,600 = 600 x 271) - so that is the layout written
here, chunk for chunk.  The ``PEAK`` chunk carries the time of writing, so no two writes of the same audio are byte-identical even
in the reference; every other byte is determined by the samples.  ``tests/test_egress.py`` pins the size and the base64 prefix of
the logged response.

No model code here, and nothing is computed on the CPU that the GPU should do: the samples arrive through the engine's pinned
staging buffer (``Engine.synthesize(audio_to_host=True)``) or as any tensor; the header is 80 bytes of host arithmetic."""
from __future__ import annotations

import base64
import struct
import time
from typing import Optional

import numpy as np
import torch

SAMPLE_RATE = 24000
HEADER_BYTES = 80


def wav_float32_bytes(audio: torch.Tensor, sample_rate: int = SAMPLE_RATE, timestamp: Optional[int] = None) -> bytes:
    """Mono 32-bit-float RIFF/WAVE file image of ``audio`` (1-D, any device / float dtype), as the reference's
    ``save_audio_chunk`` writes it to disk (service.py:373-404): ``fmt `` (16 bytes, WAVE_FORMAT_IEEE_FLOAT) + ``fact`` (sample
    count) + ``PEAK`` (version 1, time stamp, peak value and its position) + ``data``."""
    a = audio.detach().reshape(-1).to(torch.float32)
    x = a.cpu().numpy() if a.device.type != "cpu" else a.numpy()
    x = np.ascontiguousarray(x, dtype="<f4")
    n = int(x.shape[0])
    if n:
        pos = int(np.argmax(np.abs(x)))
        peak = float(abs(x[pos]))
    else:
        pos, peak = 0, 0.0
    ts = int(time.time()) if timestamp is None else int(timestamp)
    data_bytes = 4 * n
    fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 3, 1, sample_rate, sample_rate * 4, 4, 32)
    fact = struct.pack("<4sII", b"fact", 4, n)
    peak_chunk = struct.pack("<4sIIIfI", b"PEAK", 16, 1, ts & 0xFFFFFFFF, peak, pos)
    data_hdr = struct.pack("<4sI", b"data", data_bytes)
    body = b"WAVE" + fmt + fact + peak_chunk + data_hdr
    head = struct.pack("<4sI", b"RIFF", len(body) + data_bytes)
    out = head + body + x.tobytes()
    assert len(out) == HEADER_BYTES + data_bytes
    return out


def wav_base64(audio: torch.Tensor, sample_rate: int = SAMPLE_RATE, timestamp: Optional[int] = None) -> str:
    """what the endpoint puts into ``audio_base64`` (api/endpoints.py:150-158)"""
    return base64.b64encode(wav_float32_bytes(audio, sample_rate, timestamp)).decode("ascii")


def read_wav_float32(blob: bytes):
    """parse an image written by :func:`wav_float32_bytes` (tests; chunk walk, so other writers' float WAVs parse too)"""
    if blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE image")
    off, rate, data = 12, None, None
    while off + 8 <= len(blob):
        cid, size = struct.unpack_from("<4sI", blob, off)
        body = blob[off + 8: off + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag != 3 or ch != 1 or bits != 32:
                raise ValueError(f"not mono float32 (tag {tag}, channels {ch}, bits {bits})")
        elif cid == b"data":
            data = np.frombuffer(body, dtype="<f4").copy()
        off += 8 + size + (size & 1)
    if rate is None or data is None:
        raise ValueError("fmt or data chunk missing")
    return torch.from_numpy(data), rate
