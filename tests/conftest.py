import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: full-depth parity run (minutes of CPU oracle time)")


@pytest.fixture(scope="session")
def fluxlib():
    """Build (if stale) and load libfluxb200.so."""
    from diffusion_rs_b200 import build, lib

    build.build()
    return lib.load()
