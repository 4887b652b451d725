import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def mel_tables():
    """Dense fb [1025,256] + window [2048] from the reference's torchaudio objects (tests/golden/mel_fb.npz)."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "mel_fb.npz"))
    fb = np.zeros((1025, 256), np.float32)
    off = 0
    for m, (s, l) in enumerate(zip(z["start"], z["length"])):
        fb[s:s + l, m] = z["weights"][off:off + l]
        off += l
    return fb, z["window"]
