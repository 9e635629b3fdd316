"""Task table + continuous batcher in front of ``BatchKModel.forward_batch``.

Mirror of the reference's ``TTSServiceManager`` (``/root/reference/src/illufly_tts/core/service.py``): same task states
(:22-27), same method names and return shapes - ``submit_task`` (:72-116), ``cancel_task`` (:118-135), ``get_task_status``
(:137-160), ``get_user_tasks`` (:162-176), ``stream_result`` (:178-219), ``start`` (:221-229) - with a different batching
policy behind them.

What the reference's loop does (:244-271): every round it takes the FIRST pending task of each user, ordered by
``sequence_id``, until ``batch_size`` (4) tasks are collected.  One user with a hundred queued requests therefore runs at
batch size 1, and nobody runs above 4, while the engine is ~34x more productive at 32 utterances than at 1
(profiles/r01_configs.md).  What that policy encodes, and what is kept here:
  * per-user order: a user's tasks complete in ``sequence_id`` order, never a later one before an earlier one;
  * fairness: batches are filled round-robin over users, so a heavy user cannot starve a light one.
What changes: a batch is filled up to a TOKEN BUDGET (the engine's workspace is proportional to it), it may hold several
tasks of one user and all chunks of a long text (config 3b: 2.6x over running the chunks one after another), the loop
sleeps on an event instead of polling every 100 ms (:250), and requests that arrive while a batch runs form the next batch.

Deliberate deviations:
  * the model should be built with ``padded_lstm=False``: the reference's ``forward_batch`` lets the padding of the longest
    utterance leak into the durations of the others, so a request's audio would depend on who else was
    queued.  Per-utterance semantics make ``pred_dur`` of a request bit-identical to synthesising it alone;
  * a failing batch is bisected and retried, and only the requests that still fail are FAILED (the reference fails all of
    them, :336-343; at 32 per batch one bad request would take 31 others with it).  The model must RAISE on failure
    (``BatchKModel(strict=True)``): the non-strict class returns silence for the whole batch, like kmodel.py:147-150;
  * ``stream_result`` also waits while the task is PENDING (the reference falls through and yields nothing, :187-189).

Out of scope here as in §8: text normalisation / G2P (``g2p`` is a callable the caller provides; default: the text already
is one phoneme string), wav files, HTTP.  There is no CPU fallback: the model is the CUDA engine.
"""
from __future__ import annotations

import asyncio
import logging
import time
import uuid
from enum import Enum
from typing import Any, AsyncGenerator, Callable, Dict, List, Mapping, Optional, Sequence

import torch

logger = logging.getLogger(__name__)


class TaskStatus(Enum):                       # service.py:22-27
    PENDING = "pending"
    PROCESSING = "processing"
    COMPLETED = "completed"
    CANCELED = "canceled"
    FAILED = "failed"


class TTSTask:                                # service.py:29-44
    def __init__(self, task_id: str, text: str, voice_id: str, speed: float = 1.0, user_id: Optional[str] = None):
        self.task_id = task_id
        self.text = text
        self.voice_id = voice_id
        self.speed = speed
        self.user_id = user_id
        self.status = TaskStatus.PENDING
        self.created_at = time.time()
        self.started_at: Optional[float] = None
        self.completed_at: Optional[float] = None
        self.result = None
        self.error: Optional[str] = None
        self.audio_chunks: List[torch.Tensor] = []
        self.debug_id: Optional[str] = None
        self.sequence_id: float = time.time()
        self.chunks: List[str] = []           # phoneme strings of the text's chunks (one utterance each)

    @property
    def cost(self) -> int:                    # tokens incl. BOS/EOS: what the engine's workspace scales with
        return sum(len(c) + 2 for c in self.chunks)


def voice_row(pack: torch.Tensor, n_symbols: int) -> torch.Tensor:
    """Style row of a voice pack [510,1,256] for an utterance of ``n_symbols`` phonemes (pipeline.py:596-601)."""
    return pack[min(max(n_symbols, 1) - 1, pack.shape[0] - 1)]


class TTSServiceManager:
    def __init__(self, model, voices: Mapping[str, torch.Tensor], *, token_budget: int = 3200, max_batch_items: int = 64,
                 max_wait_time: float = 0.0, g2p: Optional[Callable[[str], Sequence[str]]] = None,
                 load_voice: Optional[Callable[[str], torch.Tensor]] = None, output_dir: Optional[str] = None):
        """``model.forward_batch(list[str], ref_s [B,1,256], speeds) -> list[Tensor]`` is ``BatchKModel`` (or a stand-in in
        tests).  ``voices``: voice id -> pack [510,1,256]; ``load_voice(voice_id)`` (the reference's ``pipeline.load_voice``,
        pipeline.py:78-109, which fills its ``voices`` dict on demand and raises ValueError for an unknown id) is called for an id
        that is not in ``voices`` yet, as the reference's ``submit_task`` does (:88-101).  ``token_budget``: tokens (phonemes + BOS/EOS) per engine call;
        3200 is the size of BASELINE configs[1] (32 utterances of 60-90 symbols, ~5 GB of workspace).  ``max_wait_time``:
        how long a batch smaller than the budget waits for more arrivals (0 = never; the reference has the parameter,
        :53, and never reads it)."""
        if token_budget <= 0 or max_batch_items <= 0:
            raise ValueError("token_budget and max_batch_items must be positive")
        # A BatchKModel in the reference's compatibility mode (strict=False) swallows engine failures and returns a second of silence
        # per item: the failed requests would be COMPLETED with silence and the bisection below would never run.  The service needs
        # failures to raise, so it switches the model to strict; and it warns when the model batches with the reference's padded
        # duration LSTM, where a request's durations depend on the longest request it happens to share a batch with.
        if hasattr(model, "strict") and not model.strict:
            logger.warning("TTSServiceManager: switching the model to strict=True (engine failures must raise to be isolated)")
            model.strict = True
        if getattr(model, "padded_lstm", False):
            logger.warning("TTSServiceManager: the model runs with padded_lstm=True: a request's durations depend on its batch; "
                           "construct BatchKModel(..., padded_lstm=False) for batch-invariant results")
        self.task_ttl = 3600.0                             # finished tasks are forgotten after this many seconds (evict_finished)
        # The reference writes every finished task to <output_dir>/<task_id>.wav (service.py:318-321, 373-404) and its endpoint
        # re-reads that file (api/endpoints.py:125-150).  With an output_dir the same file appears here (same float32 RIFF layout,
        # egress.py) so that endpoint keeps working unchanged; without one nothing touches the disk and task_wav_bytes() /
        # task_wav_base64() give the endpoint's payload straight from memory.
        self.output_dir = output_dir
        self.model = model
        self.voices = voices
        # voice packs resident on the device when the model offers it (BatchKModel on a B200); a stand-in model gets stacked rows
        self._resident = callable(getattr(model, "register_voice", None)) and callable(getattr(model, "forward_batch_voices", None))
        self._resident_ids: Dict[str, Any] = {}
        self.token_budget = token_budget
        self.max_batch_items = max_batch_items
        self.max_wait_time = max_wait_time
        self.g2p = g2p or (lambda text: [text])
        self.load_voice = load_voice
        self.tasks: Dict[str, TTSTask] = {}
        self.processing_task: Optional[asyncio.Task] = None
        self._wake: Optional[asyncio.Event] = None
        self.batch_log: List[Dict[str, Any]] = []          # one entry per engine call: sizes only (tests, measurements)

    # ---- the reference's public surface -------------------------------------------------------------------
    async def submit_task(self, text: str, voice_id: str, speed: float = 1.0, user_id: Optional[str] = None,
                          debug_id: Optional[str] = None, sequence_id: Optional[float] = None) -> str:
        # Everything that can suspend or fail happens BEFORE the task is registered: the batch loop runs whenever this coroutine
        # awaits, and must never see a half-built PENDING task (the reference, too, loads the voice first: service.py:88-107).
        problem = None
        if voice_id not in self.voices and self.load_voice is not None:
            try:                                            # service.py:91: load on demand, off the event loop
                pack = await asyncio.to_thread(self.load_voice, voice_id)
                if voice_id not in self.voices:
                    self.voices[voice_id] = pack            # pipeline.load_voice stores it itself when `voices` is its dict
            except Exception as exc:
                problem = f"voice {voice_id!r} could not be loaded: {exc}"
        if problem is None and voice_id not in self.voices:  # service.py:88-101: a task that is born FAILED
            problem = f"voice {voice_id!r} is not loaded"
        chunks: List[str] = []
        if problem is None:
            try:
                chunks = [c for c in self.g2p(text) if c]
                if not chunks:
                    problem = "empty text"
            except Exception as exc:
                problem = f"g2p failed: {exc}"
        task_id = str(uuid.uuid4())
        task = TTSTask(task_id, text, voice_id, speed, user_id)
        task.debug_id = debug_id
        if sequence_id is not None:
            task.sequence_id = sequence_id
        task.chunks = chunks
        if problem:
            task.status = TaskStatus.FAILED
            task.error = problem
            task.completed_at = time.time()
        self.tasks[task_id] = task                          # no await between here and the return
        if not problem and self._wake is not None:
            self._wake.set()
        return task_id

    async def cancel_task(self, task_id: str) -> bool:     # only a task that has not started can be cancelled
        task = self.tasks.get(task_id)
        if task is None or task.status != TaskStatus.PENDING:
            return False
        task.status = TaskStatus.CANCELED
        return True

    async def get_task_status(self, task_id: str) -> Optional[Dict[str, Any]]:
        task = self.tasks.get(task_id)
        if task is None:
            return None
        return {"task_id": task.task_id, "status": task.status.value, "created_at": task.created_at, "started_at": task.started_at,
                "completed_at": task.completed_at, "text": task.text, "voice_id": task.voice_id, "user_id": task.user_id,
                "chunks_completed": len(task.audio_chunks)}

    async def get_user_tasks(self, user_id: str) -> List[Dict[str, Any]]:
        return [await self.get_task_status(tid) for tid, t in self.tasks.items() if t.user_id == user_id]

    async def stream_result(self, task_id: str) -> AsyncGenerator[torch.Tensor, None]:
        task = self.tasks.get(task_id)
        if task is None:
            raise ValueError(f"no such task: {task_id}")
        while task.status in (TaskStatus.PENDING, TaskStatus.PROCESSING):
            await asyncio.sleep(0.002)
        if task.status == TaskStatus.FAILED:
            raise ValueError(f"task failed: {task.error}")
        if task.status == TaskStatus.CANCELED:
            raise ValueError("task was canceled")
        for chunk in task.audio_chunks:
            yield chunk

    async def start(self) -> None:
        if self.processing_task is None or self.processing_task.done():
            self._wake = asyncio.Event()
            self._wake.set()
            self.processing_task = asyncio.create_task(self._batch_processing_loop())

    async def stop(self) -> None:
        if self.processing_task is not None:
            self.processing_task.cancel()
            try:
                await self.processing_task
            except asyncio.CancelledError:
                pass
            self.processing_task = None

    # ---- batching policy ----------------------------------------------------------------------------------
    def select_batch(self) -> List[TTSTask]:
        """Round-robin over users by depth: first the oldest pending task of every user, then the second oldest of every
        user, ... until the token budget or the item limit is reached.  A user's k-th task is only taken if the (k-1)-th was
        (order within a user); a task larger than the budget gets a batch of its own."""
        per_user: Dict[str, List[TTSTask]] = {}
        for t in self.tasks.values():
            if t.status == TaskStatus.PENDING and t.chunks:
                per_user.setdefault(t.user_id or "anonymous", []).append(t)
        for lst in per_user.values():
            lst.sort(key=lambda t: (t.sequence_id, t.created_at))
        users = sorted(per_user, key=lambda u: (per_user[u][0].sequence_id, per_user[u][0].created_at))
        batch: List[TTSTask] = []
        tokens = items = 0
        open_users = list(users)
        depth = 0
        while open_users:
            still_open = []
            for u in open_users:
                lst = per_user[u]
                if depth >= len(lst):
                    continue
                t = lst[depth]
                fits = tokens + t.cost <= self.token_budget and items + len(t.chunks) <= self.max_batch_items
                if fits or not batch:
                    batch.append(t)
                    tokens += t.cost
                    items += len(t.chunks)
                    still_open.append(u)
                # a user whose task did not fit is closed for this batch: later tasks must not overtake it
            open_users = still_open
            depth += 1
            if tokens >= self.token_budget or items >= self.max_batch_items:
                break
        return batch

    @classmethod
    def from_reference(cls, repo_id: str, voices_dir: str, device: Optional[str] = None, batch_size: int = 4, max_wait_time: float = 0.2,
                       chunk_size: int = 200, output_dir: Optional[str] = None, **kw) -> "TTSServiceManager":
        """The reference's own constructor signature (core/service.py:47-56).  Builds the reference's ``CachedTTSPipeline`` - its text
        normalisation, G2P and voice loading, unchanged - with the B200 engine installed in the ``BatchKModel`` slot
        (kmodel.install_into_reference), and puts this manager's batcher in front of it.  ``batch_size`` and ``chunk_size`` are the
        reference's knobs for ITS batcher (at most 4 tasks per call, never read respectively); the token budget replaces them
        (``token_budget=`` in ``kw``)."""
        from .kmodel import install_into_reference
        install_into_reference()
        from illufly_tts.core.pipeline import CachedTTSPipeline            # the reference's, resolved with the stand-in model class
        pipeline = CachedTTSPipeline(repo_id, voices_dir, device)

        def g2p(text: str):                                                # what pipeline.process(segment_text=True) does per segment (:400-413)
            norm = pipeline.preprocess_text(text)
            return [pipeline.phonemes_to_ipa(pipeline.text_to_phonemes(seg)) for seg in pipeline.segment_text(norm)]
        mgr = cls(pipeline.model, pipeline.voices, g2p=g2p, load_voice=pipeline.load_voice, max_wait_time=0.0, output_dir=output_dir, **kw)
        mgr.pipeline = pipeline
        mgr.batch_size, mgr.chunk_size, mgr.reference_max_wait_time = batch_size, chunk_size, max_wait_time
        return mgr

    # ---- audio egress ------------------------------------------------------------------------
    def task_audio(self, task_id: str) -> torch.Tensor:
        """the finished task's waveform: its chunks concatenated (pipeline.py:420), on the host"""
        t = self.tasks[task_id]
        if t.status != TaskStatus.COMPLETED or not t.audio_chunks:
            raise ValueError(f"task {task_id} has no audio (status {t.status.value})")
        return torch.cat([c.reshape(-1).cpu() for c in t.audio_chunks])

    def task_wav_bytes(self, task_id: str) -> bytes:
        """the file the reference would have written for this task (service.py:373-404), in memory"""
        from .egress import wav_float32_bytes
        return wav_float32_bytes(self.task_audio(task_id))

    def task_wav_base64(self, task_id: str) -> str:
        """the ``audio_base64`` field of the reference's endpoint response (api/endpoints.py:150-158)"""
        from .egress import wav_base64
        return wav_base64(self.task_audio(task_id))

    def _write_output(self, t: TTSTask) -> None:
        import os
        from .egress import wav_float32_bytes
        os.makedirs(self.output_dir, exist_ok=True)
        blob = wav_float32_bytes(torch.cat([c.reshape(-1).cpu() for c in t.audio_chunks]))
        with open(os.path.join(self.output_dir, f"{t.task_id}.wav"), "wb") as f:
            f.write(blob)

    def evict_finished(self, now: Optional[float] = None) -> int:
        """forget COMPLETED / FAILED / CANCELED tasks older than ``task_ttl`` (the reference keeps every task for ever); called
        once per batch round, so ``tasks`` - which select_batch scans - stays proportional to the live load"""
        now = time.time() if now is None else now
        dead = [k for k, t in self.tasks.items()
                if t.status not in (TaskStatus.PENDING, TaskStatus.PROCESSING) and now - (t.completed_at or t.created_at) > self.task_ttl]
        for k in dead:
            del self.tasks[k]
        return len(dead)

    def _run(self, units: List[tuple]) -> List[torch.Tensor]:
        """One engine call for ``units`` = [(phonemes, style row [1,256], speed[, voice id])].  Worker thread.
        A model with resident voice packs (``BatchKModel.register_voice`` / ``forward_batch_voices``) gets voice ids:
        the packs were uploaded once and the rows are picked on the device; any other model gets the stacked rows as the reference's
        does (pipeline.py:603)."""
        ph = [u[0] for u in units]
        if self._resident and all(len(u) > 3 and u[3] in self._resident_ids for u in units):
            out = self.model.forward_batch_voices(ph, [u[3] for u in units], [u[2] for u in units])
        else:
            ref = torch.stack([u[1] for u in units])       # [B,1,256] like pipeline.py:603
            out = self.model.forward_batch(ph, ref, [u[2] for u in units])
        if len(out) != len(units):
            raise RuntimeError(f"model returned {len(out)} waveforms for {len(units)} utterances")
        return list(out)

    def _run_isolating(self, units: List[tuple]) -> List[Any]:
        """Like ``_run``, but a failing call is bisected: the result holds a waveform or the exception per utterance."""
        try:
            t0 = time.perf_counter()
            out = self._run(units)
            self.batch_log.append({"utterances": len(units), "tokens": sum(len(u[0]) + 2 for u in units), "seconds": time.perf_counter() - t0})
            return out
        except Exception as exc:
            if len(units) <= 1:                             # also terminates on an empty list
                return [exc] * len(units)
            mid = len(units) // 2
            return self._run_isolating(units[:mid]) + self._run_isolating(units[mid:])

    async def _batch_processing_loop(self) -> None:
        assert self._wake is not None
        while True:
            self.evict_finished()
            batch = self.select_batch()
            if not batch:
                self._wake.clear()
                if any(t.status == TaskStatus.PENDING for t in self.tasks.values()):
                    continue                                # submitted between the scan and the clear
                await self._wake.wait()
                continue
            if self.max_wait_time > 0 and sum(t.cost for t in batch) < self.token_budget:
                await asyncio.sleep(self.max_wait_time)
                batch = self.select_batch()
            now = time.time()
            units: List[tuple] = []
            live: List[TTSTask] = []
            for t in batch:                                 # a task whose inputs cannot be assembled fails alone
                t.status = TaskStatus.PROCESSING
                t.started_at = now
                try:
                    pack = self.voices[t.voice_id]
                    if self._resident:
                        if self._resident_ids.get(t.voice_id) is not pack:
                            await asyncio.to_thread(self.model.register_voice, t.voice_id, pack)  # once per voice (again if the pack object changes)
                            self._resident_ids[t.voice_id] = pack
                        else:
                            self.model.register_voice(t.voice_id, pack)    # a no-op, unless the model has been moved to a new engine since
                    mine = [(c, voice_row(pack, len(c)), float(t.speed), t.voice_id) for c in t.chunks]
                except Exception as exc:
                    t.status, t.error, t.completed_at = TaskStatus.FAILED, f"could not assemble the request: {exc!r}", time.time()
                    continue
                units += mine
                live.append(t)
            if not live:
                continue
            try:
                results = await asyncio.to_thread(self._run_isolating, units)
            except Exception as exc:                        # nothing may kill the loop: every later task would hang (service.py:336-343)
                results = [exc] * len(units)
            pos = 0
            for t in live:                                  # in selection order: sequence order within every user
                mine = results[pos:pos + len(t.chunks)]
                pos += len(t.chunks)
                bad = next((r for r in mine if isinstance(r, Exception)), None)
                if bad is not None or len(mine) != len(t.chunks):
                    t.status = TaskStatus.FAILED
                    t.error = str(bad) if bad is not None else "the engine returned too few waveforms"
                else:
                    t.audio_chunks.extend(mine)
                    if self.output_dir:                     # the file the reference's endpoint re-reads: on disk BEFORE the task shows as done
                        try:
                            await asyncio.to_thread(self._write_output, t)
                        except Exception as exc:
                            t.status = TaskStatus.FAILED
                            t.error = f"could not write the audio file: {exc}"
                    if t.status != TaskStatus.FAILED:
                        t.status = TaskStatus.COMPLETED
                t.completed_at = time.time()


class ReferencePolicyManager(TTSServiceManager):
    """The reference's selection (service.py:244-271) behind the same engine and the same loop, for measurements: one
    task per user per batch, at most ``batch_size`` tasks."""

    def __init__(self, model, voices, *, batch_size: int = 4, **kw):
        super().__init__(model, voices, **kw)
        self.batch_size = batch_size

    def select_batch(self) -> List[TTSTask]:
        per_user: Dict[str, List[TTSTask]] = {}
        for t in self.tasks.values():
            if t.status == TaskStatus.PENDING:
                per_user.setdefault(t.user_id or "anonymous", []).append(t)
        batch: List[TTSTask] = []
        for lst in per_user.values():
            lst.sort(key=lambda t: t.sequence_id)
            batch.append(lst[0])
            if len(batch) >= self.batch_size:
                break
        return batch
