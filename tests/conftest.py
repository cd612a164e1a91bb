import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
for pth in (REPO, REPO / "oracle", REPO / "tests"):
    if str(pth) not in sys.path:
        sys.path.insert(0, str(pth))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")
