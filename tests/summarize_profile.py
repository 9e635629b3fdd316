"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tests/summarize_profile.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md "<command>"
    python tests/summarize_profile.py full gpurun_out/umma_r01.ncu-rep profiles/r01_conv_umma_full.md"""
import collections
import csv
import re
import subprocess
import sys


def kname(s):
    m = re.search(r"(\w+_kernel)", s)
    return m.group(1) if m else s[:40]


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
    else:
        full(sys.argv[2], sys.argv[3])
