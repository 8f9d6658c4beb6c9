import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ltc_tables():
    """Small generated GGX LTC tables (res 32, 11 Fresnel layers), quantised like load_ltc_table."""
    from risltc_b200 import ltc_fit
    fits = ltc_fit.fit_ggx_ltc(32, 11, 32)
    rgba, rg = ltc_fit.quantize_fits(fits)
    return fits, rgba, rg


@pytest.fixture(scope="session")
def device():
    from risltc_b200 import api
    dev = api.Device(0)
    yield dev
    dev.close()
