"""Oracle-side view of the synthetic data.  TEST INFRASTRUCTURE.

The seeded checkpoint / voice packs / workloads themselves are plain data generators and live in
``illufly-tts_b200/synthetic.py`` (no model code runs there); this module only loads that checkpoint into the CPU oracle."""
from __future__ import annotations

import warnings

import illufly_tts_b200.synthetic as _syn
from illufly_tts_b200.synthetic import (random_ids, synthetic_checkpoint, synthetic_vocab, synthetic_voice_pack,  # noqa: F401
                                        vocoder_inputs, workload)

from .kokoro_oracle import KokoroConfig, KokoroOracle


def build_oracle(seed: int = 1234) -> KokoroOracle:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = KokoroOracle(KokoroConfig()).eval()
    return m.load_checkpoint(_syn.synthetic_checkpoint(seed))
