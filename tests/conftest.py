import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity contract (BASELINE.json north_star): waveform within max-abs 2e-3 and >= 40 dB SNR of the
# fp32 reference forward; integer resampling/padding indices bit-exact.
MAX_ABS_TOL = 2e-3
SNR_DB_MIN = 40.0


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_sd(name):
    return {k: torch.from_numpy(v.copy()) for k, v in golden(name).items()}


@pytest.fixture(scope="session")
def hsv():
    import megatts2_hierspeechpp_b200 as pkg
    from megatts2_hierspeechpp_b200 import build
    build.build()
    return pkg
