import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# am_mode 3 pairs two AM steps per data pass only when the pass is long enough to pay for it (0.65 GB per rank); the tests use small
# problems and must exercise the pairing: no threshold (inherited by the torchrun subprocesses of the multi-GPU tests)
os.environ.setdefault("GAMC_AM_PAIR_MIN_BYTES", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gamc():
    import gamc_b200
    return gamc_b200


@pytest.fixture(scope="session")
def engine(gamc):
    eng = gamc.Engine(0)
    yield eng
    eng.close()
