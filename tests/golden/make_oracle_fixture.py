"""Regenerates tests/golden/oracle_outputs.json: what the CPU oracle (oracle/kokoro_oracle.py, seeded synthetic checkpoint 1234)
produces on BASELINE configs[0] and on the first three utterances of configs[1].
    python tests/golden/make_oracle_fixture.py
NOT a reference pin: kokoro and its weights cannot run here, so these vectors come from this repo's restatement of the reference
graph (parity stays "unpinned by the reference").  What the fixture does pin: (1) the oracle against itself
across rounds and machines (tests/test_oracle.py recomputes it), (2) the engine's integer durations against committed numbers
without an oracle run in the loop (tests/test_gpu_configs.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from illufly_tts_b200.synthetic import workload  # noqa: E402
from oracle.synth_weights import build_oracle  # noqa: E402

TAPS = ("d", "t_en", "F0", "N", "gen_in")


def cases():
    w1 = workload("batch32")
    return {"configs0_single": workload("single"),
            "configs1_first3": dict(ids=w1["ids"][:3], ref_s=w1["ref_s"][:3], speeds=w1["speeds"][:3])}


def summarise(o):
    from tests.parity import oracle_taps_cat
    out = dict(pred_dur=[[int(v) for v in p.tolist()] for p in o.pred_dur], frames=[int(p.sum()) for p in o.pred_dur], taps={}, audio=[])
    for k in TAPS:
        t = oracle_taps_cat(o, k).double()
        out["taps"][k] = dict(shape=list(t.shape), mean_abs=float(t.abs().mean()), max_abs=float(t.abs().max()))
    for a in o.audio:
        a = a.double()
        out["audio"].append(dict(samples=int(a.numel()), mean_abs=float(a.abs().mean()), max_abs=float(a.abs().max())))
    return out


def compute():
    torch.manual_seed(0)
    oracle = build_oracle(1234)
    res = {}
    for name, w in cases().items():
        o = oracle.forward_batch(w["ids"], w["ref_s"], w["speeds"], noise_seed=0, want_taps=True)
        res[name] = summarise(o)
    return res


if __name__ == "__main__":
    res = dict(generator="tests/golden/make_oracle_fixture.py", torch=torch.__version__, cases=compute())
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_outputs.json")
    json.dump(res, open(dst, "w"), indent=1)
    for k, v in res["cases"].items():
        print(k, v["frames"], {t: round(s["mean_abs"], 6) for t, s in v["taps"].items()}, [round(a["mean_abs"], 6) for a in v["audio"]])
