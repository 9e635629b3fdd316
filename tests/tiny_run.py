"""smallest end-to-end synthesis in both precisions (the command compute-sanitizer wraps)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import illufly_tts_b200 as pkg
from illufly_tts_b200.synthetic import synthetic_checkpoint, random_ids, synthetic_voice_pack
g = torch.Generator().manual_seed(3)
ids = [random_ids(g, 6), random_ids(g, 11)]
ref = synthetic_voice_pack(1)[[5, 10], 0]
eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP32)
for prec in (pkg.PRECISION_FP32, pkg.PRECISION_FP16):
    eng.set_precision(prec)
    r = eng.synthesize(ids, ref, [1.0, 1.1], seed=1, want_indices=True)
    torch.cuda.synchronize()
    print(prec, [a.numel() for a in r.audio], [float(a.abs().max()) for a in r.audio], flush=True)
eng.close()
