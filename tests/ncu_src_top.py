"""Top stall sites of one kernel launch from `ncu --page source --csv` (SASS view), gzip or plain:
    python tests/ncu_src_top.py gpurun_out/prof8_src_1.csv.gz [N]
Prints the N instructions with the most warp-stall samples, their dominant stall reasons and a few instructions of context."""
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
print(rows[0][1][:120])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for h in stall_cols:
    agg[h] = sum(int(r[ci[h]] or 0) for r in data)
print("stall mix:", {k: f"{100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[ci["# Samples"]] or 0)
    reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    ctx = " | ".join(data[j][ci["Source"]].strip()[:38] for j in range(max(0, i - 2), i))
    print(f"{i:6d} {100 * s / tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {[(h, n) for n, h in reasons if n]}   <- {ctx}")
