"""accuracy of the conv kernels against an fp64 reference (GPU): fp32 FFMA vs split-fp16 tcgen05 vs fp16 tcgen05"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import illufly_tts_b200 as pkg
from illufly_tts_b200.synthetic import synthetic_checkpoint
eng = pkg.Engine(synthetic_checkpoint(1234), 0)
g = torch.Generator().manual_seed(0)
for (Cin, Cout, k, mean) in [(768, 768, 1, 0.0), (2048, 768, 1, 0.0), (768, 2048, 1, 0.5), (640, 2048, 1, 0.0), (512, 512, 5, 0.0), (128, 768, 1, 0.0)]:
    rows = 1024
    x = torch.randn(rows, Cin, generator=g) + mean
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.conv1d(x.double().t()[None], w.double(), b.double(), padding=(k - 1) // 2)[0].t()
    scale = ref.abs().mean().item()
    out = []
    for prec, name in [(0, "fp32-ffma"), (3, "split-f16"), (2, "f16")]:
        y = eng.op_conv1d(x.cuda(), w, b, [0, rows], 1, prec).cpu().double()
        e = (y - ref)
        out.append(f"{name}: max {e.abs().max().item() / scale:.2e} rms {e.pow(2).mean().sqrt().item() / scale:.2e} bias {e.mean().item() / scale:+.2e}")
    print(f"K={Cin * k} N={Cout}: " + " | ".join(out), flush=True)
