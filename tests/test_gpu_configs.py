"""GPU parity on BASELINE.json's OWN configs at full size (-m gpu), every precision mode, through the C ABI.

Each test runs the split protocol of tests/parity.py (front half: integers bit-exact + taps; back half: the oracle's generator
on the engine's own generator inputs) AND records the plain end-to-end comparison (engine waveform against the oracle's waveform,
nothing substituted).  Measured numbers go to the parity report (tests/parity.py::record -> profiles/r02_parity.md).

The frame-side tap bars below are the values MEASURED on B200 (profiles/r02_parity.md) times two."""
import os

import pytest
import torch

import illufly_tts_b200 as pkg
from illufly_tts_b200.synthetic import vocoder_inputs, workload
from tests.parity import back_half, e2e_records, front_half, record, snr_db, summarise, vocoder_check

pytestmark = pytest.mark.gpu

MODES = {"fp32": pkg.PRECISION_FP32, "f16": pkg.PRECISION_FP16, "bf16": pkg.PRECISION_BF16, "fp32_ffma": pkg.PRECISION_FP32_FFMA}
# relative error (max-abs / max-abs of the oracle's tap) allowed on the frame-side taps, per mode: measured worst case x 2
# (measured worst case over the configs: fp32 1.5e-5 / 8.1e-6 / 1.7e-5; f16 1.6e-3 / 1.8e-3 / 2.1e-3; bf16 1.2e-2 / 8.9e-3 / 1.6e-2)
FRAME_TOL = {
    "fp32": dict(F0=3e-5, N=2e-5, gen_in=4e-5),          # the strict tier: split-fp16 operands on tcgen05
    "fp32_ffma": dict(F0=3e-5, N=2e-5, gen_in=4e-5),
    "f16": dict(F0=3.2e-3, N=3.6e-3, gen_in=4.2e-3),
    "bf16": dict(F0=2.4e-2, N=1.8e-2, gen_in=3.2e-2),
}
# back half (bit-identical generator inputs): the north star's bars - max-abs 1e-3 in fp32 (measured 2.9e-6), 40 dB in the 16-bit
# operand modes (measured: f16 59.1 dB, bf16 42.8 dB worst case)
BACK = {"fp32": dict(tol_abs=1e-3), "fp32_ffma": dict(tol_abs=1e-3), "f16": dict(min_snr=40.0), "bf16": dict(min_snr=40.0)}

_oracle_runs = {}


def oracle_run(oracle, name, w):
    """one oracle pass per workload serves every precision mode (a batch-32 pass is ~15 s of CPU)"""
    if name not in _oracle_runs:
        _oracle_runs[name] = oracle.forward_batch(w["ids"], w["ref_s"], w["speeds"], noise_seed=0, want_taps=True,
                                                  tap_keys=("d", "t_en", "en", "asr", "F0", "N", "gen_in"))
    return _oracle_runs[name]


def run_case(engine, oracle, name, w, mode):
    engine.set_precision(MODES[mode])
    try:
        fh = front_half(engine, oracle, w, frame_tol=FRAME_TOL[mode], o=oracle_run(oracle, name, w))
        back = back_half(engine, oracle, w, fh, **BACK[mode])
        e2e = e2e_records(fh)
        record(name, mode, utterances=len(w["ids"]), tokens=sum(len(x) for x in w["ids"]), frames=sum(fh["Fs"]),
               front={k: fh["report"][k] for k in ("duration_abs", "duration_margin", "d", "t_en", "en", "asr", "F0", "N", "gen_in")},
               back=summarise(back), e2e=summarise(e2e))
        return fh, back, e2e
    finally:
        engine.set_precision(pkg.PRECISION_FP32)


@pytest.mark.parametrize("mode", ["fp32", "fp32_ffma", "f16", "bf16"])
def test_config0_single(engine, oracle, mode):
    """BASELINE configs[0]: the single short utterance"""
    run_case(engine, oracle, "configs[0] single", workload("single"), mode)


@pytest.mark.parametrize("mode", ["fp32", "f16", "bf16"])
def test_config1_batch32_full(engine, oracle, mode):
    """BASELINE configs[1] = the bench workload, all 32 utterances at full size: durations and the alignment map bit-exact, every
    tap within its bar, the waveform of every utterance within the north star's bar on bit-identical generator inputs, and the
    end-to-end waveform comparison recorded"""
    fh, back, e2e = run_case(engine, oracle, "configs[1] batch32", workload("batch32"), mode)
    assert len(back) == 32 and sum(r["n"] for r in back) == 600 * sum(fh["Fs"])


@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_config2_paragraph_chunk(engine, oracle, mode):
    """BASELINE configs[2]: one ~350-symbol chunk of the long paragraph (B=1, the pipeline.py:400-420 path)"""
    w = workload("paragraph")
    k = min(range(6), key=lambda i: abs(len(w["ids"][i]) - 352))
    w1 = dict(ids=[w["ids"][k]], ref_s=w["ref_s"][k:k + 1], speeds=[1.0])
    assert 300 <= len(w1["ids"][0]) <= 402
    run_case(engine, oracle, "configs[2] chunk", w1, mode)


@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_config4_fleet_sample(engine, oracle, mode):
    """BASELINE configs[4]: a 16-utterance sample of the fleet (16 voice packs, speeds 0.8-1.2, 40-120 symbols), one batch"""
    w = workload("fleet16")
    assert len({tuple(r.tolist()[:4]) for r in w["ref_s"]}) == 16 and min(w["speeds"]) < 0.95 < 1.05 < max(w["speeds"])
    run_case(engine, oracle, "configs[4] fleet16", w, mode)


@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_config3_vocoder_cell_b16_l1000(engine, oracle, mode):
    """BASELINE configs[3]: one large cell of the vocoder sweep (B=16, L=1000 generator frames = 12.5 s each)"""
    B, L = 16, 1000
    x, s, f0, noise = vocoder_inputs(B, L)
    engine.set_precision(MODES[mode])
    engine.enable_taps(True)
    try:
        audio = engine.vocoder(x, s, f0, noise)
        har = engine.read_tap("har")
    finally:
        engine.enable_taps(False)
        engine.set_precision(pkg.PRECISION_FP32)
    n = 60 * L + 1
    back, e2e = [], []
    for b in range(B):
        tol = BACK[mode]
        back.append(vocoder_check(oracle, x[b:b + 1], s[b:b + 1], f0[b:b + 1], noise[b:b + 1], har[b * n:(b + 1) * n], audio[b],
                                  tol.get("tol_abs"), tol.get("min_snr"), f"B={B} L={L} item {b}"))
        if b < 4:      # end to end (the oracle's own source spectrum, nothing substituted): four items are enough for the record
            ref = oracle.vocoder(x[b:b + 1], s[b:b + 1], f0[b:b + 1], noise[b:b + 1])[0]
            a = audio[b].cpu()
            e2e.append(dict(max_abs=float((a - ref).abs().max()), snr=snr_db(ref, a), n=a.numel()))
    record("configs[3] vocoder B=16 L=1000", mode, utterances=B, frames=B * L, back=summarise(back), e2e=summarise(e2e))


@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_engine_durations_equal_the_committed_vectors(engine, mode):
    """tests/golden/oracle_outputs.json: integer durations (and so frame and sample counts) of configs[0] and of three utterances of
    configs[1], committed - the engine must reproduce them with no oracle run in the loop, in the strict and in the 16-bit mode (the
    text side is fp32-class in both)."""
    import importlib.util
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(here, "oracle_outputs.json")))["cases"]
    spec = importlib.util.spec_from_file_location("make_oracle_fixture", os.path.join(here, "make_oracle_fixture.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    engine.set_precision(MODES[mode])
    try:
        for name, w in mod.cases().items():
            r = engine.synthesize(w["ids"], w["ref_s"], w["speeds"], seed=1)
            assert [[int(v) for v in p.tolist()] for p in r.pred_dur] == gold[name]["pred_dur"], name
            assert [int(a.numel()) for a in r.audio] == [600 * f for f in gold[name]["frames"]], name
    finally:
        engine.set_precision(pkg.PRECISION_FP32)
