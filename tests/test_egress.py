"""Audio egress: the in-memory float32 RIFF/WAVE image against the one byte-level fact the reference tree
holds - the response logged at /root/reference/service.log:33 (650,480 bytes, base64 prefix ``UklGRujsCQ``) - and round trips."""
import base64
import json
import os
import struct

import torch

from illufly_tts_b200.egress import HEADER_BYTES, read_wav_float32, wav_base64, wav_float32_bytes


GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_service_log.json")))


def test_size_and_prefix_of_the_logged_reference_response():
    """service.log:33 shows a RIFF of 650,480 bytes whose base64 starts with UklGRujsCQ (tests/golden/reference_service_log.json,
    extracted from the reference tree by tests/golden/make_reference_log_fixture.py).  162,600 samples (= 600 x 271 frames, the only
    float-WAV header size that leaves a whole number of 600-sample frames) written here give exactly that size and prefix."""
    assert GOLDEN["audio_base64_prefix"] == "UklGRujsCQ" and GOLDEN["file_bytes"] == 650_480        # the fixture is what the survey read
    n = (GOLDEN["file_bytes"] - HEADER_BYTES) // 4
    assert n == 600 * 271 and HEADER_BYTES == 80
    audio = torch.zeros(n)
    blob = wav_float32_bytes(audio, 24000, timestamp=0)
    assert len(blob) == GOLDEN["file_bytes"]
    assert struct.unpack_from("<I", blob, 4)[0] == GOLDEN["riff_size_field"] == 0x0009ECE8
    assert blob[:7].hex() == GOLDEN["decoded_hex"]
    assert base64.b64encode(blob).decode()[:10] == GOLDEN["audio_base64_prefix"]
    assert wav_base64(audio, 24000, timestamp=0)[:10] == GOLDEN["audio_base64_prefix"]


def test_header_fields_and_round_trip():
    g = torch.Generator().manual_seed(0)
    audio = torch.randn(600 * 7, generator=g) * 0.3
    audio[1234] = -1.75
    blob = wav_float32_bytes(audio, 24000, timestamp=1_700_000_000)
    assert blob[:4] == b"RIFF" and blob[8:16] == b"WAVEfmt "
    size, tag, ch, rate, byterate, align, bits = struct.unpack_from("<IHHIIHH", blob, 16)
    assert (size, tag, ch, rate, byterate, align, bits) == (16, 3, 1, 24000, 96000, 4, 32)          # WAVE_FORMAT_IEEE_FLOAT, mono
    assert blob[36:40] == b"fact" and struct.unpack_from("<II", blob, 40) == (4, audio.numel())
    assert blob[48:52] == b"PEAK"
    psize, version, ts, peak, pos = struct.unpack_from("<IIIfI", blob, 52)
    assert (psize, version, ts, pos) == (16, 1, 1_700_000_000, 1234) and abs(peak - 1.75) < 1e-6
    assert blob[72:76] == b"data" and struct.unpack_from("<I", blob, 76)[0] == 4 * audio.numel()
    back, sr = read_wav_float32(blob)
    assert sr == 24000 and torch.equal(back, audio)                                              # the samples are carried bit for bit
    # every byte but the PEAK time stamp is a function of the samples
    other = wav_float32_bytes(audio, 24000, timestamp=5)
    assert blob[:60] == other[:60] and blob[64:] == other[64:] and blob[60:64] != other[60:64]


def test_dtype_device_and_empty():
    a = torch.linspace(-1, 1, 1200, dtype=torch.float64)
    back, _ = read_wav_float32(wav_float32_bytes(a))
    assert torch.equal(back, a.float())
    assert len(wav_float32_bytes(torch.zeros(0))) == HEADER_BYTES
