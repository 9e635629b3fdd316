"""Host logic of the continuous batcher (illufly-tts_b200/service.py) against the properties the reference's service encodes
(/root/reference/src/illufly_tts/core/service.py): task states and status dict, cancel only while pending, per-user order,
fairness across users; plus what the new policy adds: token budget, several chunks per task, failure isolation.
CPU only: the model is a recording stand-in with the ``forward_batch`` signature of BatchKModel."""
import asyncio
import threading
import time

import pytest
import torch

import illufly_tts_b200 as pkg
from illufly_tts_b200.service import voice_row

VOICES = {"v1": torch.arange(510 * 256, dtype=torch.float32).reshape(510, 1, 256), "v2": torch.ones(510, 1, 256)}
STATUS_KEYS = {"task_id", "status", "created_at", "started_at", "completed_at", "text", "voice_id", "user_id", "chunks_completed"}   # service.py:150-160


class FakeModel:
    """forward_batch(list[str], ref_s [B,1,256], speeds) -> list of 1-D tensors; remembers every call."""

    def __init__(self, delay=0.0, poison=None, gate: threading.Event = None):
        self.calls, self.delay, self.poison, self.gate = [], delay, poison, gate

    def forward_batch(self, phonemes, ref_s, speeds=None):
        assert ref_s.shape == (len(phonemes), 1, 256)
        self.calls.append(list(phonemes))
        if self.gate is not None:
            self.gate.wait(5.0)
        if self.delay:
            time.sleep(self.delay)
        if self.poison is not None and any(self.poison in p for p in phonemes):
            raise RuntimeError("poisoned batch")
        return [torch.full((600 * len(p),), float(len(p)) + float(ref_s[i, 0, 0]) * 1e-6) for i, p in enumerate(phonemes)]


def run(coro):
    return asyncio.run(coro)


async def drain(svc, ids, timeout=10.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        st = [svc.tasks[i].status for i in ids]
        if all(s in (pkg.TaskStatus.COMPLETED, pkg.TaskStatus.FAILED, pkg.TaskStatus.CANCELED) for s in st):
            return
        await asyncio.sleep(0.002)
    raise AssertionError("tasks did not finish")


def test_lifecycle_and_status_dict():
    async def main():
        svc = pkg.TTSServiceManager(FakeModel(), VOICES)
        await svc.start()
        tid = await svc.submit_task("abcdefgh", "v1", 1.1, user_id="u", debug_id="d", sequence_id=3.0)
        st = await svc.get_task_status(tid)
        assert set(st) == STATUS_KEYS and st["status"] in ("pending", "processing", "completed") and st["text"] == "abcdefgh"
        await drain(svc, [tid])
        st = await svc.get_task_status(tid)
        assert st["status"] == "completed" and st["chunks_completed"] == 1 and st["started_at"] <= st["completed_at"]
        chunks = [c async for c in svc.stream_result(tid)]
        assert len(chunks) == 1 and chunks[0].numel() == 600 * 8
        assert await svc.get_task_status("nope") is None
        assert [s["task_id"] for s in await svc.get_user_tasks("u")] == [tid]
        await svc.stop()
    run(main())


def test_unknown_voice_is_born_failed_and_stream_raises():
    async def main():
        svc = pkg.TTSServiceManager(FakeModel(), VOICES)
        await svc.start()
        tid = await svc.submit_task("abcdef", "nobody")
        st = await svc.get_task_status(tid)
        assert st["status"] == "failed" and st["completed_at"] is not None                 # service.py:88-101
        with pytest.raises(ValueError):
            [c async for c in svc.stream_result(tid)]
        with pytest.raises(ValueError):
            [c async for c in svc.stream_result("nope")]
        await svc.stop()
    run(main())


def test_cancel_only_while_pending():
    async def main():
        gate = threading.Event()
        model = FakeModel(gate=gate)
        svc = pkg.TTSServiceManager(model, VOICES, token_budget=12)
        a = await svc.submit_task("aaaaaaaaaa", "v1", user_id="u")        # 12 tokens: fills the first batch alone
        b = await svc.submit_task("bbbbbbbbbb", "v1", user_id="u")
        await svc.start()
        while not model.calls:
            await asyncio.sleep(0.001)
        assert svc.tasks[a].status == pkg.TaskStatus.PROCESSING
        assert not await svc.cancel_task(a)                                # service.py:133-135: already started
        assert await svc.cancel_task(b) and not await svc.cancel_task(b) and not await svc.cancel_task("nope")
        gate.set()
        await drain(svc, [a, b])
        assert svc.tasks[a].status == pkg.TaskStatus.COMPLETED and svc.tasks[b].status == pkg.TaskStatus.CANCELED
        assert model.calls == [["aaaaaaaaaa"]]                             # the canceled task never reached the engine
        with pytest.raises(ValueError):
            [c async for c in svc.stream_result(b)]
        await svc.stop()
    run(main())


def test_budget_fairness_and_per_user_order():
    async def main():
        model = FakeModel(delay=0.002)
        svc = pkg.TTSServiceManager(model, VOICES, token_budget=50, max_batch_items=64)
        heavy = [await svc.submit_task("h" * 8 + f"{i:02d}", "v1", user_id="heavy", sequence_id=float(i)) for i in range(20)]    # 12 tokens each
        light = await svc.submit_task("l" * 10, "v2", user_id="light", sequence_id=100.0)
        await svc.start()
        await drain(svc, heavy + [light])
        assert any("l" * 10 in call for call in model.calls[:1]), "the light user's only task must ride in the first batch"
        for call in model.calls:
            assert sum(len(p) + 2 for p in call) <= 50
        assert max(len(c) for c in model.calls) == 4 and len(model.calls) < 21     # batches are filled, not one task per user
        done = [svc.tasks[t].completed_at for t in heavy]
        assert done == sorted(done), "a user's tasks complete in sequence order"
        order = [p for call in model.calls for p in call if p.startswith("h")]
        assert order == sorted(order), "and reach the engine in sequence order"
        await svc.stop()
    run(main())


def test_later_task_never_overtakes_one_that_does_not_fit():
    svc = pkg.TTSServiceManager(FakeModel(), VOICES, token_budget=30)

    async def main():
        await svc.submit_task("a" * 10, "v1", user_id="u", sequence_id=1.0)      # 12
        await svc.submit_task("b" * 25, "v1", user_id="u", sequence_id=2.0)      # 27: does not fit after the first
        await svc.submit_task("c" * 4, "v1", user_id="u", sequence_id=3.0)       # 6: would fit, but must wait for b
        await svc.submit_task("d" * 4, "v1", user_id="w", sequence_id=4.0)
    run(main())
    assert [t.text[0] for t in svc.select_batch()] == ["a", "d"]


def test_oversize_task_gets_its_own_batch_and_chunks_stay_together():
    async def main():
        model = FakeModel()
        svc = pkg.TTSServiceManager(model, VOICES, token_budget=20, g2p=lambda text: text.split("|"))
        big = await svc.submit_task("x" * 30 + "|" + "y" * 30 + "|" + "z" * 5, "v1", user_id="a")
        small = await svc.submit_task("s" * 6, "v1", user_id="b")
        await svc.start()
        await drain(svc, [big, small])
        assert model.calls[0] == ["x" * 30, "y" * 30, "z" * 5] and model.calls[1] == ["s" * 6]
        chunks = [c async for c in svc.stream_result(big)]
        assert [c.numel() for c in chunks] == [600 * 30, 600 * 30, 600 * 5]          # chunk order of the text
        assert (await svc.get_task_status(big))["chunks_completed"] == 3
        await svc.stop()
    run(main())


def test_voice_row_follows_length_like_the_pipeline():
    pack = VOICES["v1"]
    assert torch.equal(voice_row(pack, 1), pack[0]) and torch.equal(voice_row(pack, 75), pack[74]) and torch.equal(voice_row(pack, 9999), pack[509])

    async def main():
        seen = []

        class M(FakeModel):
            def forward_batch(self, phonemes, ref_s, speeds=None):
                seen.append((list(phonemes), ref_s.clone(), list(speeds)))
                return super().forward_batch(phonemes, ref_s, speeds)
        svc = pkg.TTSServiceManager(M(), VOICES)
        t = await svc.submit_task("q" * 7, "v1", 0.9)
        await svc.start()
        await drain(svc, [t])
        ph, ref, sp = seen[0]
        assert torch.equal(ref[0], pack[6]) and sp == [0.9]
        await svc.stop()
    run(main())


def test_failure_is_isolated_by_bisection():
    async def main():
        model = FakeModel(poison="BAD")
        svc = pkg.TTSServiceManager(model, VOICES, token_budget=1000)
        ids = [await svc.submit_task(("BAD" if i == 5 else "ok") + "p" * 6 + str(i), "v1", user_id=f"u{i}") for i in range(9)]
        await svc.start()
        await drain(svc, ids)
        st = [svc.tasks[i].status for i in ids]
        assert st[5] == pkg.TaskStatus.FAILED and "poisoned" in svc.tasks[ids[5]].error
        assert all(s == pkg.TaskStatus.COMPLETED for k, s in enumerate(st) if k != 5), "one bad request must not fail its neighbours"
        assert len(model.calls[0]) == 9 and len(model.calls) <= 1 + 2 * 4             # one full call, then a bisection
        await svc.stop()
    run(main())


def test_reference_policy_is_one_task_per_user_at_most_four():
    async def main():
        model = FakeModel()
        svc = pkg.ReferencePolicyManager(model, VOICES, batch_size=4)
        ids = [await svc.submit_task("p" * 6 + f"{u}{i}", "v1", user_id=f"u{u}", sequence_id=float(i)) for u in range(6) for i in range(3)]
        single = [await svc.submit_task("s" * 6 + str(i), "v1", user_id="solo", sequence_id=10.0 + i) for i in range(3)]
        await svc.start()
        await drain(svc, ids + single)
        for call in model.calls:
            assert len(call) <= 4 and len({p[6] for p in call if p.startswith("p")}) == len([p for p in call if p.startswith("p")])
        await svc.stop()
    run(main())


def test_voices_are_loaded_on_demand_like_the_reference():
    """pipeline.voices starts empty and pipeline.load_voice fills it (pipeline.py:78-109); submit_task loads before it judges
    (service.py:88-101): a voice that is merely not loaded yet must not fail the task, an unknown one must."""
    async def main():
        store, loads = {}, []

        def load_voice(vid):
            loads.append(vid)
            if vid not in VOICES:
                raise ValueError(f"no such voice {vid}")
            store[vid] = VOICES[vid]                       # the reference's loader stores into its own dict
            return VOICES[vid]
        svc = pkg.TTSServiceManager(FakeModel(), store, load_voice=load_voice)
        await svc.start()
        a = await svc.submit_task("abcdefg", "v1")
        b = await svc.submit_task("abcdefg", "v1")
        c = await svc.submit_task("abcdefg", "ghost")
        await drain(svc, [a, b, c])
        assert svc.tasks[a].status == svc.tasks[b].status == pkg.TaskStatus.COMPLETED
        assert svc.tasks[c].status == pkg.TaskStatus.FAILED and "ghost" in svc.tasks[c].error
        assert loads == ["v1", "ghost"], "a loaded voice is not loaded again"
        await svc.stop()
    run(main())


def test_slow_voice_load_never_exposes_a_half_built_task():
    """submit_task suspends while the voice loads; the batch loop runs meanwhile and must not pick the task up before its
    chunks exist (regression: a PENDING task without chunks gave the engine an empty batch)."""
    async def main():
        store = {"v2": VOICES["v2"]}

        def slow_load(vid):
            time.sleep(0.05)
            store[vid] = VOICES[vid]
            return VOICES[vid]
        model = FakeModel()
        svc = pkg.TTSServiceManager(model, store, load_voice=slow_load)
        await svc.start()
        slow = asyncio.create_task(svc.submit_task("slowslow", "v1", user_id="a"))
        await asyncio.sleep(0.005)
        fast = await svc.submit_task("fastfast", "v2", user_id="b")        # wakes the loop while the other submit is suspended
        await drain(svc, [fast])
        assert all(len(c) > 0 for c in model.calls)
        tid = await slow
        await drain(svc, [tid])
        assert svc.tasks[tid].status == pkg.TaskStatus.COMPLETED and all(len(c) > 0 for c in model.calls)
        assert [c.numel() for c in svc.tasks[tid].audio_chunks] == [600 * 8], "the slow task must receive its audio"
        assert ["slowslow"] in model.calls
        await svc.stop()
    run(main())


def test_bisection_terminates_on_degenerate_input():
    svc = pkg.TTSServiceManager(FakeModel(poison="x"), VOICES)
    assert svc._run_isolating([]) == []
    out = svc._run_isolating([("x" * 6, VOICES["v1"][0], 1.0)])
    assert len(out) == 1 and isinstance(out[0], Exception)


def test_a_task_that_cannot_be_assembled_fails_alone_and_the_loop_survives():
    """the loop must outlive anything a single request can do to it (the reference guards its batch body, service.py:336-343)"""
    async def main():
        store = dict(VOICES)
        gate = threading.Event()
        svc = pkg.TTSServiceManager(FakeModel(gate=gate), store, token_budget=12)
        first = await svc.submit_task("a" * 10, "v1", user_id="u1")          # fills batch 1 alone and blocks on the gate
        await svc.start()
        while svc.tasks[first].status != pkg.TaskStatus.PROCESSING:
            await asyncio.sleep(0.001)
        doomed = await svc.submit_task("b" * 10, "v2", user_id="u2")
        okay = await svc.submit_task("c" * 10, "v1", user_id="u3", speed="not a number")
        fine = await svc.submit_task("d" * 10, "v1", user_id="u4")
        del store["v2"]                                                       # the voice disappears before its batch is built
        gate.set()
        await drain(svc, [first, doomed, okay, fine])
        assert svc.tasks[doomed].status == pkg.TaskStatus.FAILED and "assemble" in svc.tasks[doomed].error
        assert svc.tasks[okay].status == pkg.TaskStatus.FAILED
        assert svc.tasks[first].status == svc.tasks[fine].status == pkg.TaskStatus.COMPLETED
        later = await svc.submit_task("e" * 10, "v1", user_id="u5")          # the loop is still alive
        await drain(svc, [later])
        assert svc.tasks[later].status == pkg.TaskStatus.COMPLETED and not svc.processing_task.done()
        await svc.stop()
    run(main())


def test_service_forces_strict_and_evicts_finished_tasks():
    """a model in the reference's compatibility mode (failures -> silence) would make failed requests COMPLETED with silence: the
    service switches it to strict; finished tasks are forgotten after task_ttl"""
    import asyncio

    class Compat:
        strict = False
        padded_lstm = False

        def forward_batch(self, ph, ref, speeds=None):
            return [torch.zeros(600 * len(p)) for p in ph]
    m = Compat()
    svc = pkg.TTSServiceManager(m, {"v": torch.zeros(510, 1, 256)})
    assert m.strict is True

    async def main():
        tid = await svc.submit_task("abcdefgh", "v")
        await svc.start()
        async for _ in svc.stream_result(tid):
            pass
        await svc.stop()
        return tid
    tid = asyncio.run(main())
    assert svc.tasks[tid].status == pkg.TaskStatus.COMPLETED
    assert svc.evict_finished(now=svc.tasks[tid].completed_at + 10) == 0 and tid in svc.tasks
    assert svc.evict_finished(now=svc.tasks[tid].completed_at + svc.task_ttl + 1) == 1 and tid not in svc.tasks


def test_service_audio_egress_and_reference_constructor(tmp_path, monkeypatch):
    """SURVEY 8(f) row 2 + the reference's constructor signature (service.py:47-56): from_reference() builds the reference's
    pipeline class (here a stand-in module with the methods the manager uses) around the installed BatchKModel slot; a finished task
    is available as the float32 RIFF image the reference's endpoint returns - from memory, and as <output_dir>/<task_id>.wav"""
    import asyncio
    import sys
    import types

    from illufly_tts_b200.egress import read_wav_float32

    class FakeModel:
        strict = True
        padded_lstm = False

        def forward_batch(self, ph, ref, speeds=None):
            return [torch.full((600 * len(p),), 0.01 * len(p)) for p in ph]

    class FakePipeline:                                   # the surface of the reference's CachedTTSPipeline that the manager touches
        def __init__(self, repo_id, voices_dir, device=None):
            self.model, self.voices, self.args = FakeModel(), {}, (repo_id, voices_dir, device)

        def load_voice(self, vid):
            self.voices[vid] = torch.zeros(510, 1, 256)
            return self.voices[vid]

        def preprocess_text(self, text): return text.strip()
        def segment_text(self, text): return [s for s in text.split("|") if s]
        def text_to_phonemes(self, seg): return seg.lower()
        def phonemes_to_ipa(self, ph): return ph + "ipa"
    mod = types.ModuleType("illufly_tts.core.pipeline")
    mod.CachedTTSPipeline = FakePipeline
    for name in ("illufly_tts", "illufly_tts.core"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, "illufly_tts.core.pipeline", mod)

    async def main():
        svc = pkg.TTSServiceManager.from_reference("repo", "voices", "cuda:0", batch_size=4, max_wait_time=0.2, chunk_size=200, output_dir=str(tmp_path))
        assert svc.pipeline.args == ("repo", "voices", "cuda:0") and sys.modules["illufly_tts.core.kmodel"].BatchKModel is pkg.BatchKModel
        tid = await svc.submit_task("Hello wor|ld again", "zf_001")
        await svc.start()
        async for _ in svc.stream_result(tid):
            pass
        await svc.stop()
        return svc, tid
    svc, tid = asyncio.run(main())
    t = svc.tasks[tid]
    assert t.status == pkg.TaskStatus.COMPLETED and t.chunks == ["hello woripa", "ld againipa"]
    blob = svc.task_wav_bytes(tid)
    audio, sr = read_wav_float32(blob)
    assert sr == 24000 and audio.numel() == 600 * (12 + 11) and torch.equal(audio, svc.task_audio(tid))
    on_disk = (tmp_path / f"{tid}.wav").read_bytes()
    assert on_disk[:60] == blob[:60] and on_disk[64:] == blob[64:]           # identical but for the PEAK time stamp
    import base64
    assert base64.b64decode(svc.task_wav_base64(tid))[:4] == b"RIFF"
    sys.modules.pop("illufly_tts.core.kmodel", None)


class ResidentModel(FakeModel):
    """a model with device-resident voice packs (BatchKModel.register_voice / forward_batch_voices): remembers what it was given"""

    def __init__(self):
        super().__init__()
        self.registered, self.voice_calls = [], []
        self.packs = {}

    def register_voice(self, voice_id, pack):
        if self.packs.get(voice_id) is pack:               # BatchKModel.register_voice: the same tensor object again is a no-op
            return list(self.packs).index(voice_id)
        self.registered.append(voice_id)
        self.packs[voice_id] = pack
        return len(self.packs) - 1

    def forward_batch_voices(self, phonemes, voice_ids, speeds=None):
        self.voice_calls.append((list(phonemes), list(voice_ids)))
        rows = torch.stack([voice_row(self.packs[v], len(p)) for p, v in zip(phonemes, voice_ids)])
        return [torch.full((600 * len(p),), float(len(p)) + float(rows[i, 0, 0]) * 1e-6) for i, p in enumerate(phonemes)]


def test_resident_voice_packs_are_registered_once_and_selected_by_id():
    """with a model that keeps voice packs on the device the service uploads a pack once per voice and then sends
    voice ids, not stacked style rows (pipeline.py:596-603 stacks rows on the host for every batch); the waveforms equal what the
    stacked-row path gives.  A stand-in model without the two methods keeps the reference's path (every other test of this file)."""
    async def main(model):
        svc = pkg.TTSServiceManager(model, dict(VOICES), token_budget=40)
        await svc.start()
        ids = [await svc.submit_task("abcdefgh" + "x" * k, "v1" if k % 2 else "v2", 1.0, user_id=f"u{k % 3}") for k in range(7)]
        await drain(svc, ids)
        out = [svc.tasks[i].audio_chunks[0] for i in ids]
        await svc.stop()
        return out
    res, plain = ResidentModel(), FakeModel()
    a = run(main(res))
    b = run(main(plain))
    assert sorted(res.registered) == ["v1", "v2"]                      # once per voice, not once per batch
    assert res.voice_calls and not res.calls                           # only the resident entry point was used
    assert all(set(v) <= {"v1", "v2"} for _, v in res.voice_calls)
    assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
