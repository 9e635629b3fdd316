"""CPU tests (-m "not gpu") of the host side: the C-ABI library loads and exports every symbol the header declares, the
BatchKModel mirror keeps the reference's surface and error behaviour, partitioning, and the N>1 sharding over gloo."""
import os
import re
import subprocess
import sys

import pytest
import torch

import illufly_tts_b200 as pkg
from illufly_tts_b200 import _lib
from illufly_tts_b200.kmodel import phonemes_to_ids
from illufly_tts_b200.synthetic import ids_to_phonemes, synthetic_checkpoint, synthetic_config, synthetic_vocab, workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "ttsb200.h")).read()
    declared = set(re.findall(r"TTSB_API\s+[\w\s\*]+?\b(ttsb_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                      # dlopen works without a GPU; no compute call is made here
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ttsb_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ttsb_\w+)", out))
    assert exported == declared


def test_engine_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.EngineError) as ei:
        pkg.Engine(synthetic_checkpoint(1234), 0)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_no_cpu_device():
    m = pkg.BatchKModel(config=synthetic_config(), model={k: {} for k in ("bert", "bert_encoder", "predictor", "decoder", "text_encoder")})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.to("cpu")
    with pytest.raises(RuntimeError, match="call .to"):
        m.engine


def test_phonemes_to_ids_matches_reference_mapping():
    vocab = synthetic_vocab()
    assert len(vocab) == 177 and min(vocab.values()) == 1 and max(vocab.values()) == 177
    s = "ab世c"                     # unknown character dropped (kmodel.py:38-39), BOS/EOS = 0 (kmodel.py:41)
    assert phonemes_to_ids(vocab, s) == [0, vocab["a"], vocab["b"], vocab["c"], 0]
    ids = workload("single")["ids"][0]
    assert phonemes_to_ids(vocab, ids_to_phonemes(ids)) == ids


def test_forward_batch_error_surface_without_engine():
    """kmodel.py:27-30: any phoneme string shorter than 5 -> ONE-element list of zeros(24000), before any compute;
    kmodel.py:147-150: any failure -> B x zeros(24000)"""
    m = pkg.BatchKModel(config=synthetic_config(), model={k: {} for k in ("bert", "bert_encoder", "predictor", "decoder", "text_encoder")})
    out = m.forward_batch(["abcdefgh", "abc"], torch.zeros(2, 1, 256))
    assert len(out) == 1 and out[0].shape == (24000,) and float(out[0].abs().sum()) == 0.0
    out = m.forward_batch(["abcdefgh", "abcdefg"], torch.zeros(2, 1, 256))          # no engine -> exception -> silence per item
    assert len(out) == 2 and all(o.shape == (24000,) for o in out)
    strict = pkg.BatchKModel(config=synthetic_config(), model={k: {} for k in ("bert", "bert_encoder", "predictor", "decoder", "text_encoder")}, strict=True)
    with pytest.raises(ValueError):
        strict.forward_batch(["abc"], torch.zeros(1, 1, 256))
    with pytest.raises(RuntimeError):
        strict.forward_batch(["abcdefgh"], torch.zeros(1, 1, 256))


def test_install_into_reference_module_slot():
    mod = pkg.install_into_reference()
    assert sys.modules["illufly_tts.core.kmodel"].BatchKModel is pkg.BatchKModel
    assert mod.BatchKModel is pkg.BatchKModel
    del sys.modules["illufly_tts.core.kmodel"]


def test_checkpoint_manifest_matches_generator(checkpoint):
    from illufly_tts_b200.synthetic import manifest
    man = manifest()
    for mod, sd in checkpoint.items():
        assert {k[7:] for k in sd} == set(man[mod])
        for k, v in sd.items():
            assert list(v.shape) == man[mod][k[7:]], (mod, k)
    ck2 = synthetic_checkpoint(1234)
    assert torch.equal(ck2["decoder"]["module.generator.ups.1.weight_v"], checkpoint["decoder"]["module.generator.ups.1.weight_v"])


def test_lpt_partition_and_micro_batches():
    costs = [5, 9, 2, 7, 7, 1, 3, 8]
    parts = pkg.lpt_partition(costs, 3)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert pkg.lpt_partition(costs, 3) == parts                     # deterministic
    assert pkg.lpt_partition([], 2) == [[], []]
    mbs = pkg.micro_batches(costs, 10)
    assert sorted(i for p in mbs for i in p) == list(range(8))
    assert all(sum(costs[i] for i in p) <= 10 or len(p) == 1 for p in mbs)
    assert pkg.micro_batches([50], 10) == [[0]]


def test_world_size_2_sharding_gloo(tmp_path):
    """the N>1 path: ranks partition the utterances (no collective on the data path), then max-over-ranks time and summed
    units are reduced exactly as bench.py does — exercised with gloo on CPU, world_size 2"""
    script = tmp_path / "w.py"
    script.write_text(f'''
import os, sys, json
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
import illufly_tts_b200 as pkg
from illufly_tts_b200.synthetic import workload
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w = workload("fleet64")
costs = [len(x) for x in w["ids"]]
mine = pkg.lpt_partition(costs, world)[rank]
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
units = torch.tensor([float(sum(costs[i] for i in mine)), float(len(mine))], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.all_reduce(units, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps(dict(tmax=t.item(), tokens=units[0].item(), n=units[1].item(), want=float(sum(costs)))))
dist.barrier()
dist.destroy_process_group()
''')
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613",
                        str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["tmax"] == 2.0 and d["n"] == 64 and d["tokens"] == d["want"]


def test_bench_fleet_partition_world_size_2_gloo(tmp_path):
    """bench.py --workload fleet (BASELINE configs[4]): every rank builds the SAME fleet, takes its LPT share and cuts it into
    token-budgeted micro-batches; over two gloo ranks every utterance is synthesised exactly once, no micro-batch exceeds the budget,
    and the predicted work of the two shares differs by less than one utterance"""
    script = tmp_path / "f.py"
    script.write_text(f'''
import os, sys, json, importlib.util
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join({ROOT!r}, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w, mine, mbs, my_cost, all_cost = bench.make_fleet(256, rank, world, 1200)
seen = torch.zeros(256, dtype=torch.int32)
for mb in mbs:
    assert sum(len(w["ids"][i]) for i in mb) <= 1200 or len(mb) == 1
    for i in mb:
        seen[i] += 1
dist.all_reduce(seen)
cost = torch.tensor([my_cost], dtype=torch.float64)
lo, hi = cost.clone(), cost.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    longest = max(len(x) / s for x, s in zip(w["ids"], w["speeds"]))
    print(json.dumps(dict(once=bool((seen == 1).all()), spread=hi.item() - lo.item(), longest=longest, nmb=len(mbs))))
dist.barrier()
dist.destroy_process_group()
''')
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29617",
                        str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["once"] and d["spread"] <= d["longest"] and d["nmb"] >= 5


def test_bench_reference_arm_runs_on_cpu():
    """--impl reference is the CPU oracle arm: must print one JSON line without a GPU"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-utts", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_clock_sampler_nvml_and_fallbacks():
    """bench.py's clocks field: in-process NVML samples (>= 2 even for an instantaneous region), reasons decoded from the bit
    mask, device resolved by UUID first; nvidia-smi, then an explicit 'none', when NVML cannot be used"""
    import importlib.util
    import time as _t
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeNvml:
        NVML_CLOCK_SM = 1
        nvmlClocksEventReasonSwPowerCap = 0x4
        nvmlClocksEventReasonHwSlowdown = 0x8
        nvmlClocksEventReasonSwThermalSlowdown = 0x20
        nvmlClocksEventReasonHwThermalSlowdown = 0x40

        def __init__(self, bits, uuid_ok=True):
            self.bits, self.uuid_ok, self.by = bits, uuid_ok, None

        def nvmlInit(self): pass

        def nvmlDeviceGetHandleByUUID(self, u):
            if not self.uuid_ok:
                raise RuntimeError("no such uuid")
            self.by = ("uuid", u)
            return "H"

        def nvmlDeviceGetHandleByIndex(self, i):
            self.by = ("index", i)
            return "H"

        def nvmlDeviceGetClockInfo(self, h, k): return 1890
        def nvmlDeviceGetMaxClockInfo(self, h, k): return 1965
        def nvmlDeviceGetCurrentClocksEventReasons(self, h): return self.bits

    nv = FakeNvml(0x4 | 0x1)                                 # sw_power_cap + "gpu idle" (not a reason we report)
    c = bench.ClockSampler(3, nvml=nv, uuid="GPU-abc")
    c.open(); c.start(); _t.sleep(0.09); out = c.stop()
    assert out["source"] == "nvml" and nv.by == ("uuid", "GPU-abc")
    assert out["sm_mhz"] == 1890 and out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"] and out["samples"] >= 4
    c = bench.ClockSampler(3, nvml=FakeNvml(0x8 | 0x40, uuid_ok=False), uuid="GPU-abc")
    c.start(); out = c.stop()                                # start() opens on its own; a zero-length region still has 2 samples
    assert c.nvml.by == ("index", 3) and out["samples"] >= 2 and out["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]

    class Broken:
        def nvmlInit(self): raise RuntimeError("no driver")
    c = bench.ClockSampler(0, nvml=Broken())
    c.open()
    assert c.source in ("nvidia-smi", "none")                # here: no nvidia-smi either, or one that prints nothing
    out = c.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
