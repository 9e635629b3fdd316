"""Where the host time of one Engine.synthesize call goes (BASELINE configs[1]): Python before / between / after the two C calls,
the calls themselves, and the device-timed step.  Median of 10 after 4 warm-ups.      python tests/host_split.py"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import synthetic_checkpoint, workload  # noqa: E402

eng = pkg.Engine(synthetic_checkpoint(1234), 0, pkg.PRECISION_FP16)
w = workload("batch32")
lib = eng._lib
marks = {}


def wrap(name):
    f = getattr(lib, name)

    def g(*a):
        t0 = time.perf_counter()
        r = f(*a)
        marks[name] = (t0, time.perf_counter())
        return r
    setattr(lib, name, g)


wrap("ttsb_synthesize_begin")
wrap("ttsb_synthesize_finish")
rows = []
for i in range(14):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = eng.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=7)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    b0, b1 = marks["ttsb_synthesize_begin"]
    f0, f1 = marks["ttsb_synthesize_finish"]
    if i >= 4:
        rows.append(((b0 - t0) * 1e3, (b1 - b0) * 1e3, (f0 - b1) * 1e3, (f1 - f0) * 1e3, (t1 - f1) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, eng.last_step_ms()))
names = ["python before begin()", "begin() call (host prep + GPU half + wait for the durations)", "python between the calls", "finish() call",
         "python after finish()", "synchronize after synthesize returned", "TOTAL wall incl. that synchronize", "device-timed step"]
for k, n in enumerate(names):
    print(f"{n:64s} {statistics.median(r[k] for r in rows):8.3f} ms")
