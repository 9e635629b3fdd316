"""Helper of test_phase_scan_parallel_equals_serial: synthesise BASELINE configs[1] (and two edge lengths) with the taps on and save
the audio-rate harmonic source and the waveforms.      python tests/dump_source.py out.pt      (TTSB_PHASE_SERIAL=1 selects the
serial phase scan; the switch is read once per process, hence a helper process per mode)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import random_ids, synthetic_checkpoint, synthetic_voice_pack, workload  # noqa: E402

eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP32)
eng.enable_taps(True)
out = {}
w = workload("batch32")
r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=5)
out["b32_src"] = eng.read_tap("har_source").clone()
out["b32_audio"] = torch.cat([a.cpu() for a in r.audio])
g = torch.Generator().manual_seed(8)
lens = [1, 33, 510]                                    # 1 frame-ish, a partial chunk of 32, many chunks
pack = synthetic_voice_pack(1)
r = eng.synthesize([random_ids(g, n) for n in lens], torch.stack([pack[n - 1, 0] for n in lens]), [1.0, 0.8, 1.2], seed=6)
out["edge_src"] = eng.read_tap("har_source").clone()
out["edge_audio"] = torch.cat([a.cpu() for a in r.audio])
torch.save(out, sys.argv[1])
print("saved", {k: tuple(v.shape) for k, v in out.items()})
