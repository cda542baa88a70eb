import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nominal_model():
    from spi_active_b200 import go2_model as gm
    return gm.go2_nominal()


@pytest.fixture(scope="session")
def blob(nominal_model):
    from spi_active_b200 import go2_model as gm
    return gm.build_model_blob(nominal_model)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session", params=["ws", "lane"])
def engine(request):
    """One engine per rollout kernel: the warp-specialised fast path and the generic leg-per-lane kernel run the
    whole GPU suite."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spi_active_b200.engine import RolloutEngine
    eng = RolloutEngine()
    eng.set_kernel(request.param)
    return eng
