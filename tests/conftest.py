import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle.synth_weights import build_oracle
    return build_oracle(1234)


@pytest.fixture(scope="session")
def checkpoint(oracle):
    return oracle.export_checkpoint()


@pytest.fixture(scope="session")
def engine(checkpoint):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import illufly_tts_b200 as pkg
    eng = pkg.Engine(checkpoint, 0)
    yield eng
    eng.close()
