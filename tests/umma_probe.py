"""GPU probe of the tcgen05 conv kernel against torch (bf16-rounded operands, fp32 accumulate).
    python tests/umma_probe.py            # runs itself twice: TTSB_UMMA_SWAP=0 and 1"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    import torch.nn.functional as F
    import illufly_tts_b200 as pkg
    from illufly_tts_b200.synthetic import synthetic_checkpoint
    eng = pkg.Engine(synthetic_checkpoint(1234), 0)
    g = torch.Generator().manual_seed(0)
    shapes = [(128, 128, 3, 1), (128, 128, 11, 5), (256, 256, 7, 3), (128, 24, 7, 1), (640, 1024, 3, 1), (512, 64, 1, 1), (64, 128, 3, 1)]
    for (Cin, Cout, k, d) in shapes:
        lens = [300, 77, 129]
        off = [0]
        for n in lens:
            off.append(off[-1] + n)
        x = torch.randn(off[-1], Cin, generator=g)
        w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
        b = torch.randn(Cout, generator=g)
        xb, wb = x.bfloat16().float(), w.bfloat16().float()
        refs = []
        for i, n in enumerate(lens):
            xi = xb[off[i]:off[i + 1]].t()[None]
            refs.append(F.conv1d(xi, wb, b, dilation=d, padding=(k * d - d) // 2)[0].t())
        ref = torch.cat(refs, 0)
        xd = x.cuda()
        try:
            y32 = eng.op_conv1d(xd, w, b, off, d, pkg.PRECISION_FP32).cpu()
            y16, ms = eng.op_conv1d(xd, w, b, off, d, pkg.PRECISION_BF16, reps=1, timed=True)
            y16 = y16.cpu()
            e32 = (y32 - ref).abs().max().item()
            e16 = (y16 - ref).abs().max().item()
            print(f"Cin={Cin} Cout={Cout} k={k} d={d}: fp32-vs-bf16ref {e32:.3e}  umma-vs-bf16ref {e16:.3e}  (ref max {ref.abs().max():.2f}) {ms:.3f} ms", flush=True)
        except Exception as exc:
            print(f"Cin={Cin} Cout={Cout} k={k} d={d}: FAILED {exc}", flush=True)
            break


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for swap in ("0",):
            print(f"==== TTSB_UMMA_SWAP={swap}", flush=True)
            env = dict(os.environ, TTSB_UMMA_SWAP=swap)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=150)
                print("exit", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", flush=True)
