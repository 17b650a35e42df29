import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the C-ABI library and the oracle once per session (nvcc cross-compiles without a GPU)
    from multiplexnanopore_b200 import build as smx_build
    smx_build.build()
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def engine():
    from multiplexnanopore_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
