"""Throughput of the service loop on one GPU: the continuous token-budget batcher against the reference's selection policy
(one task per user per batch, at most 4; service.py:244-271), SAME engine, SAME loop, closed-loop saturation (all requests
queued before the loop starts).  Wall clock, Python overhead of the loop included; one untimed pass of the same workload,
then the median of 5 timed passes with their spread.
    python tests/bench_service.py [n_requests] [probe]
`probe` also prints how well a replayed engine call and a solo call agree with what the tasks received (sets the
assertions of tests/test_gpu_parity.py::test_service_*)."""
import asyncio
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import illufly_tts_b200 as pkg  # noqa: E402
from illufly_tts_b200.synthetic import ids_to_phonemes, synthetic_checkpoint, synthetic_config, synthetic_voice_pack, workload  # noqa: E402


class Recorder:
    def __init__(self, model):
        self.model, self.calls = model, []

    def forward_batch(self, phonemes, ref_s, speeds=None):
        out = self.model.forward_batch(phonemes, ref_s, speeds)
        self.calls.append((list(phonemes), ref_s.clone(), list(speeds), [a.clone() for a in out]))
        return out


async def saturate(svc, reqs):
    ids = [await svc.submit_task(ph, v, s, user_id=u, sequence_id=float(k)) for k, (ph, v, s, u) in enumerate(reqs)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    await svc.start()
    while any(svc.tasks[i].status in (pkg.TaskStatus.PENDING, pkg.TaskStatus.PROCESSING) for i in ids):
        await asyncio.sleep(0.001)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    await svc.stop()
    assert all(svc.tasks[i].status == pkg.TaskStatus.COMPLETED for i in ids), [svc.tasks[i].error for i in ids if svc.tasks[i].error][:3]
    audio_s = sum(sum(c.numel() for c in svc.tasks[i].audio_chunks) for i in ids) / 24000.0
    return ids, audio_s, wall


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
    probe = "probe" in sys.argv
    model = pkg.BatchKModel(config=synthetic_config(), model=synthetic_checkpoint(1234), precision=pkg.PRECISION_FP16, strict=True, padded_lstm=False).to("cuda")
    voices = {f"v{k}": synthetic_voice_pack(100 + k) for k in range(16)}
    w = workload("fleet512")
    phon = [ids_to_phonemes(x) for x in w["ids"]][:n]
    reqs16 = [(phon[k], f"v{k % 16}", float(w["speeds"][k]), f"user{k % 16}") for k in range(n)]
    reqs1 = [(p, v, s, "solo") for (p, v, s, _) in reqs16]
    REPS = 5
    print(f"| policy | users | requests | engine calls | mean utterances per call | audio-s | wall s: median (min - max of {REPS}) | audio-s/s (median) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for name, mk, reqs, users in (("continuous, 3200-token budget", lambda: pkg.TTSServiceManager(model, voices, token_budget=3200), reqs16, 16),
                                  ("reference policy (1 per user, <= 4)", lambda: pkg.ReferencePolicyManager(model, voices, batch_size=4), reqs16, 16),
                                  ("continuous, 3200-token budget", lambda: pkg.TTSServiceManager(model, voices, token_budget=3200), reqs1, 1),
                                  ("reference policy (1 per user, <= 4)", lambda: pkg.ReferencePolicyManager(model, voices, batch_size=4), reqs1[:min(n, 128)], 1)):
        asyncio.run(saturate(mk(), reqs))                  # untimed pass of the SAME workload: workspace growth, first touch, thread pool
        walls = []
        for _ in range(REPS):
            svc = mk()
            ids, audio_s, wall = asyncio.run(saturate(svc, reqs))
            walls.append(wall)
        walls.sort()
        med = walls[len(walls) // 2]
        calls = len(svc.batch_log)
        print(f"| {name} | {users} | {len(reqs)} | {calls} | {sum(b['utterances'] for b in svc.batch_log) / calls:.1f} | {audio_s:.0f} | "
              f"{med:.3f} ({walls[0]:.3f} - {walls[-1]:.3f}) | {audio_s / med:.0f} |", flush=True)
    if probe:
        rec = Recorder(model)
        svc = pkg.TTSServiceManager(rec, voices, token_budget=1500)
        ids, _, _ = asyncio.run(saturate(svc, reqs16[:96]))
        got = [svc.tasks[i].audio_chunks[0] for i in ids]
        # (1) replay every recorded engine call: is the engine bit-reproducible run to run?
        worst_replay, exact = float("inf"), 0
        for ph, ref, sp, out in rec.calls:
            again = model.forward_batch(ph, ref, sp)
            for a, b in zip(out, again):
                assert a.numel() == b.numel()
                if torch.equal(a, b):
                    exact += 1
                else:
                    e = (a - b).float()
                    worst_replay = min(worst_replay, float(10 * torch.log10(a.float().pow(2).sum() / e.pow(2).sum())))
        print(f"replay: {exact} of {sum(len(c[0]) for c in rec.calls)} waveforms bit-identical; worst SNR of the others {worst_replay:.1f} dB")
        # (2) a request alone against the same request inside its service batch
        from illufly_tts_b200.service import voice_row
        same_len, snrs = 0, []
        for k in (0, 7, 33, 64, 95):
            ph, v, s, _ = reqs16[k]
            solo = model.forward_batch([ph], voice_row(voices[v], len(ph)).unsqueeze(0), [s])[0]
            same_len += int(solo.numel() == got[k].numel())
            if solo.numel() == got[k].numel():
                e = (solo - got[k]).float()
                snrs.append(float(10 * torch.log10(solo.float().pow(2).sum() / e.pow(2).sum().clamp_min(1e-30))))
        print(f"solo vs in-batch: {same_len} of 5 with identical sample count; SNR {['%.1f' % x for x in snrs]} dB")


if __name__ == "__main__":
    main()
