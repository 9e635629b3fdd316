"""Turn the ncu CSVs of tests/gpu_ncu_final.sh (brought back in gpurun_out/) into the tracked summaries under profiles/:
    python tests/summarize_ncu_r02.py
writes profiles/r02_launches.md (+ .csv copy), profiles/r02_conv_launches.md, profiles/r02_conv_wa_ncu.md, profiles/r02_traffic.json"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
CMD = "python tests/profile_step.py f16   (BASELINE configs[1], 32 utterances, second = warm step)"


def kname(s):
    m = re.search(r"(\w+_kernel)", s)
    return m.group(1) if m else s[:40]


def val(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v


def to_ms(r):
    return val(r) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]


# ---- 1. launch list ----
rows = list(csv.DictReader(l for l in open(os.path.join(G, "r02_launches.csv")) if l.startswith('"')))
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = kname(r["Kernel Name"])
    tot[k][0] += 1
    tot[k][1] += to_ms(r)
s = sum(v[1] for v in tot.values())
shutil.copy(os.path.join(G, "r02_launches.csv"), os.path.join(P, "r02_launches_batch32_f16.csv"))
with open(os.path.join(P, "r02_launches_batch32_f16.md"), "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none), end of round 2\n\ncommand: `{CMD}`\n\n")
    f.write(f"{len(rows)} launches of one step, {s:.2f} ms summed (cold-cache, serialised: compare SHARES, not absolutes; the same step timed with CUDA events: 21.5 ms)\n\n")
    f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / s:.1f} % |\n")

# ---- 2. every conv launch of the step: time, DRAM bytes, tensor pipe ----
rows = list(csv.DictReader(l for l in open(os.path.join(G, "r02_conv_launches.csv")) if l.startswith('"')))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], dict(kernel=kname(r["Kernel Name"]), name=r["Kernel Name"]))
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["ms"] = to_ms(r)
    elif m.startswith("dram__bytes_read"):
        d["rd"] = val(r) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
    elif m.startswith("dram__bytes_write"):
        d["wr"] = val(r) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
    elif m.startswith("sm__pipe_tensor"):
        d["tensor"] = val(r)
    elif m.startswith("gpu__dram_throughput"):
        d["dram_pct"] = val(r)
L = list(by_id.values())
total_ms = sum(d["ms"] for d in L)
total_bytes = sum(d.get("rd", 0) + d.get("wr", 0) for d in L)
# order of the conv launches inside a step (engine.cu): text side 0-58, shared LSTM 59, F0/N 60-73 (two streams: interleaved), decoder 74-90,
# stage 0: 91 noise_conv0, 92-97 wa noise_res0 (k7), 98 ups0, 99-116 wa resblocks (k3 x6, k7 x6, k11 x6),
# stage 1: 117 noise_conv1, 118-123 wa noise_res1 (k11), 124 ups1, 125-142 wa resblocks (k3 x6, k7 x6, k11 x6), 143 conv_post
groups = [("text side + duration head (split-fp16 GEMMs / convs)", 0, 59), ("shared LSTM proj, F0 / N blocks, decoder", 59, 91),
          ("stage 0: noise_convs[0] (144 -> 256 regrouped, k3)", 91, 92), ("stage 0: noise_res[0] k7 (weights-as-A, 256 ch)", 92, 98), ("stage 0: ups[0]", 98, 99),
          ("stage 0: resblock k3 (weights-as-A, 256 ch)", 99, 105), ("stage 0: resblock k7 (weights-as-A, 256 ch)", 105, 111),
          ("stage 0: resblock k11 (weights-as-A, 256 ch)", 111, 117), ("stage 1: noise_convs[1] (24 -> 128, k1)", 117, 118),
          ("stage 1: noise_res[1] k11 (weights-as-A)", 118, 124), ("stage 1: ups[1]", 124, 125), ("stage 1: resblock k3 (weights-as-A)", 125, 131),
          ("stage 1: resblock k7 (weights-as-A)", 131, 137), ("stage 1: resblock k11 (weights-as-A)", 137, 143), ("conv_post (128 -> 24, k7)", 143, 144)]
with open(os.path.join(P, "r02_conv_launches.md"), "w") as f:
    f.write(f"# Every convolution launch of one warm step under ncu: time, DRAM bytes, tensor-pipe share\n\ncommand: `{CMD}`, "
            "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,"
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_umma -s 144 -c 144`\n\n")
    f.write(f"{len(L)} launches, {total_ms:.2f} ms summed, {total_bytes / 1e9:.2f} GB of DRAM traffic = {total_bytes / len(L) / 1e6:.1f} MB per launch "
            "(this mean is `roofline.traffic` of bench.py). Per-launch times under ncu are cold-cache and serialised; the classes' SHARES are what carries over.\n\n")
    f.write("| launches of the step | n | ms | share | DRAM GB | DRAM throughput % (mean) | sm__pipe_tensor_cycles_active % (mean, min - max) |\n|---|---:|---:|---:|---:|---:|---|\n")
    for name, a, b in groups:
        g = L[a:b]
        if not g:
            continue
        ms = sum(d["ms"] for d in g)
        by = sum(d.get("rd", 0) + d.get("wr", 0) for d in g)
        t = [d.get("tensor", 0) for d in g]
        dr = [d.get("dram_pct", 0) for d in g]
        f.write(f"| {name} | {len(g)} | {ms:.3f} | {100 * ms / total_ms:.1f} % | {by / 1e9:.2f} | {sum(dr) / len(dr):.1f} | {sum(t) / len(t):.1f} ({min(t):.1f} - {max(t):.1f}) |\n")
    f.write("\n## Resblock launches one by one (c1 = dilated conv writing the fp16 intermediate, c2 = conv + residual)\n\n")
    f.write("| launch | kernel | ms | DRAM MB read / written | DRAM throughput % | tensor pipe % |\n|---|---|---:|---|---:|---:|\n")
    blocks = [(92, "stage 0 noise_res k7"), (99, "stage 0 k3"), (105, "stage 0 k7"), (111, "stage 0 k11"), (118, "stage 1 noise_res k11"), (125, "stage 1 k3"),
              (131, "stage 1 k7"), (137, "stage 1 k11")]
    for b0, label in blocks:
        for i in range(b0, b0 + 6):
            d = L[i]
            which = "c1" if (i - b0) % 2 == 0 else "c2"
            f.write(f"| {i} {label} {which} | `{d['kernel']}` | {d['ms']:.3f} | {d.get('rd', 0) / 1e6:.0f} / {d.get('wr', 0) / 1e6:.0f} | {d.get('dram_pct', 0):.1f} | {d.get('tensor', 0):.1f} |\n")
json.dump(dict(dram_bytes_per_launch_mean=total_bytes / len(L), launches=len(L), dram_bytes_total=total_bytes,
               source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over all 144 conv launches of one warm batch-32 step of the "
                      "final binary (profiles/r02_conv_launches.md); per launch like roofline.achieved"),
          open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)

# ---- 3. --set full of the 42 resblock launches of the weights-as-A kernel after stage 0's noise_res ----
rows = list(csv.reader(open(os.path.join(G, "r02_wa_full_raw.csv"))))
hdr, data = rows[0], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("launch__registers_per_thread", "regs"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
cols = [(c, n) for c, n in cols if c in ci]
labels = []
for st, ks in (("s0 256ch", ("k3", "k7", "k11")), ("s1 128ch", ("noise_res k11", "k3", "k7", "k11"))):
    for k in ks:
        labels += [f"{st} {k} c1", f"{st} {k} c2"] * 3
with open(os.path.join(P, "r02_conv_wa_ncu.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none --import-source on` of 42 resblock launches of `conv_umma_wa_kernel` (final binary)\n\n"
            f"command: `{CMD}`, `-k regex:conv_umma_wa -s 54 -c 42` (warm step: stage 0 resblocks k3 / k7 / k11 x6 each (256 channels), then stage 1 noise_res k11 x6 "
            "and resblocks k3 / k7 / k11 x6 each (128 channels))\n\n")
    f.write("| launch | " + " | ".join(n for _, n in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
    for i, r in enumerate(data):
        if len(r) < len(hdr) - 2:
            continue
        f.write(f"| {labels[i] if i < len(labels) else i} | " + " | ".join(r[ci[c]] for c, _ in cols) + " |\n")
    f.write("\nUnits as ncu prints them (time in us, bytes in the unit of the second CSV row: " + ", ".join(f"{n}: {rows[1][ci[c]]}" for c, n in cols if 'bytes' in c) + ").\n")
print(open(os.path.join(P, "r02_conv_launches.md")).read()[:3500])
print(open(os.path.join(P, "r02_conv_wa_ncu.md")).read())
print(open(os.path.join(P, "r02_launches_batch32_f16.md")).read())
