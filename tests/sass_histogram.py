"""SASS opcode histogram of every object of the library (the evidence that the kernels are tcgen05 / bulk-copy / cluster code):
    python tests/sass_histogram.py > profiles/r02_sass_opcodes.md
Runs `cuobjdump -sass` on illufly-tts_b200/build/*.o (built by `make -C illufly-tts_b200`); no GPU needed."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "STAS", "MUFU", "HMMA", "FFMA", "FFMA2", "DADD",
       "LDG", "STG", "LDS", "STS", "ATOM", "RED", "ACQBULK", "ERRBAR", "PLOP3", "SHFL", "BAR", "LDL", "STL"]
print("# SASS opcode histogram per object (`cuobjdump -sass illufly-tts_b200/build/*.o`, sm_100a)\n")
print("`python tests/sass_histogram.py`. Counts are static instructions over all kernels of an object. UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
print("UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (the bulk-copy / TMA engine, no tensor map), UBLKPF = cp.async.bulk.prefetch.L2,")
print("SYNCS = mbarrier operations, STAS = st.async (distributed shared memory), LDL / STL = local-memory (spill) traffic.\n")
for obj in sorted(glob.glob(os.path.join(ROOT, "illufly-tts_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels = re.findall(r"Function : (\S+)", out)
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", out, re.M):
        ops[m.group(1)] += 1
    if not kernels:
        continue
    total = sum(ops.values())
    print(f"## {os.path.basename(obj)} - {len(kernels)} kernels, {total} instructions\n")
    shown = [(k, ops[k]) for k in KEY if ops.get(k)]
    print("| " + " | ".join(k for k, _ in shown) + " |")
    print("|" + "---:|" * len(shown))
    print("| " + " | ".join(str(v) for _, v in shown) + " |\n")
    top = ", ".join(f"{k} {v}" for k, v in ops.most_common(12))
    print(f"most frequent: {top}\n")
