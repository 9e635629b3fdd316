"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tests/summarize_profile.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md "<command>"
    python tests/summarize_profile.py full gpurun_out/umma_r01.ncu-rep profiles/r01_conv_umma_full.md
    python tests/summarize_profile.py roles gpurun_out/roles.log profiles/r01_conv_umma_roles.md "<command>"   (TTSB_UMMA_PROF=1 log)"""
import collections
import csv
import re
import subprocess
import sys


def kname(s):
    m = re.search(r"(\w+_kernel)", s)
    return m.group(1) if m else s[:40]


def roles(src, dst, cmd):
    """TTSB_UMMA_PROF=1 log (stderr of tests/profile_step.py) -> per layer class: launches, CUDA-event time, busy cycles per role."""
    text = open(src).read()
    body = text.split("step 0")[1] if "step 0" in text else text          # second (warm) step only
    pat = re.compile(r"umma pro=(\d) fmt=(\d) Cin=(\d+) Cout=(\d+) k=(\d+) dil=(\d+) grid=(\d+) items/cta=([\d.]+) \| producer (\d+) \(wait (\d+)\) \| "
                     r"mma (\d+) \(wait slab (\d+) acc (\d+) w (\d+)\) \| epilogue (\d+) \(wait (\d+)\) cycles \| epilogue min (\d+) max (\d+) \| ([\d.]+) us")
    agg = collections.OrderedDict()
    tot = 0.0
    for m in pat.finditer(body):
        pro, fmt, cin, cout, k, dil, grid, items, pt, pw, mt, ms, ma, mw, et, ew, emin, emax, us = m.groups()
        key = (int(fmt), int(cin), int(cout), int(k), int(pro))
        a = agg.setdefault(key, dict(n=0, us=0.0, items=0.0, p=0, m=0, mw=0, e=0, t=0, grid=int(grid)))
        a["n"] += 1; a["us"] += float(us); a["items"] += float(items)
        a["p"] += int(pt) - int(pw); a["m"] += int(mt) - int(ms) - int(ma) - int(mw); a["mw"] += int(mw); a["e"] += int(et) - int(ew); a["t"] += int(emax)
        tot += float(us)
    with open(dst, "w") as f:
        f.write(f"# conv_umma_kernel: where the time goes, per layer class\n\ncommand: `{cmd}`\n\n")
        f.write("Cycle counters per role (`clock64`, mean over CTAs; busy = total - time blocked on mbarriers), CUDA-event time per launch;\n"
                "synchronous diagnostic mode, second step. fmt 1 = fp16 operands, 2 = split fp16 (text side); pro 0 none, 1 LeakyReLU,\n"
                "2 AdaIN+LeakyReLU, 3 AdaIN+Snake.  `k cycles / tile` columns are per 128-row x BN work item of one CTA.\n\n")
        f.write(f"{sum(a['n'] for a in agg.values())} launches, {tot / 1000:.2f} ms\n\n")
        f.write("| fmt | Cin | Cout | k | pro | launches | grid | ms | share | tile total | producer busy | MMA busy | MMA waits weights | epilogue busy |\n")
        f.write("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            it = a["items"]
            f.write(f"| {key[0]} | {key[1]} | {key[2]} | {key[3]} | {key[4]} | {a['n']} | {a['grid']} | {a['us'] / 1000:.3f} | {100 * a['us'] / tot:.1f} % | "
                    f"{a['t'] / it / 1000:.1f} k | {a['p'] / it / 1000:.1f} k | {a['m'] / it / 1000:.1f} k | {a['mw'] / it / 1000:.1f} k | {a['e'] / it / 1000:.1f} k |\n")


def launches(src, dst, cmd):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v}[r["Metric Unit"]]
        k = kname(r["Kernel Name"])
        tot[k][0] += 1
        tot[k][1] += v
    s = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n")
        f.write(f"{len(rows)} launches of one step, {s:.2f} ms summed (cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / s:.1f} % |\n")
        f.write("\n## every launch over 0.25 ms, in stream order\n\n| # | kernel | grid | ms |\n|---:|---|---|---:|\n")
        for i, r in enumerate(rows):
            v = float(r["Metric Value"].replace(",", ""))
            v = {"ns": v / 1e6, "us": v / 1e3, "ms": v}[r["Metric Unit"]]
            if v > 0.25:
                f.write(f"| {i} | {kname(r['Kernel Name'])} | {r['Grid Size']} | {v:.3f} |\n")


WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {src}\n\n| launch | kernel | " + " | ".join(w for w, _ in idx) + " |\n")
        f.write("|---|---|" + "---|" * len(idx) + "\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"| {n} | {kname(r[hdr.index('Kernel Name')])} | " + " | ".join(f"{r[i]} {units[i]}" for _, i in idx) + " |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif sys.argv[1] == "roles":
        roles(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
