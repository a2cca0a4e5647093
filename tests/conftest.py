import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    from yocto_hair_b200 import _capi
    return _capi.load()


@pytest.fixture(scope="session")
def ref():
    """The reference build (oracle/_ref, kind 'reference') when it travelled with
    the snapshot, otherwise the plain-C restatement (kind 'port')."""
    import oracle
    return oracle.best_oracle()


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.PortOracle()
