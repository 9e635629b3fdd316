"""illufly-tts B200 engine: a from-scratch sm_100a implementation of the synthesis hot path behind
``BatchKModel`` (/root/reference/src/illufly_tts/core/kmodel.py:5-150).

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/ttsb200.h) -> lib/libttsb200.so
  _lib.py      ctypes binding of the C ABI (no torch types cross it)
  engine.py    thin host wrapper: checkpoint hand-over, ragged batch descriptors, device buffers
  kmodel.py    drop-in mirror of the reference's ``BatchKModel`` interface
  partition.py utterance partitioning across GPUs / micro-batches (no collective on the data path)

There is no CPU fallback: importing works anywhere, creating an engine needs the built library and a B200.
"""
from .engine import Engine, SynthesisResult, PRECISION_FP32, PRECISION_BF16, PRECISION_FP16, PRECISION_FP32_FFMA  # noqa: F401
from .kmodel import BatchKModel, install_into_reference  # noqa: F401
from .egress import wav_base64, wav_float32_bytes  # noqa: F401
from .partition import lpt_partition, micro_batches  # noqa: F401
from .service import ReferencePolicyManager, TaskStatus, TTSServiceManager, TTSTask  # noqa: F401

__all__ = ["Engine", "SynthesisResult", "BatchKModel", "install_into_reference", "lpt_partition", "micro_batches", "wav_float32_bytes", "wav_base64", "TTSServiceManager", "ReferencePolicyManager", "TaskStatus", "TTSTask",
           "PRECISION_FP32", "PRECISION_BF16", "PRECISION_FP16", "PRECISION_FP32_FFMA"]
