"""Shared fixtures. `backend` runs every golden test against the CPU oracle (always) and the
device engine (marked gpu): same frontend, same assertions — that is the parity contract."""
import subprocess
import sys
from pathlib import Path

import pytest

from tests._pkg import ORACLE_LIB, ROOT, pkg  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _oracle_api():
    if not ORACLE_LIB.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True)
    return pkg.CApi(ORACLE_LIB, "cxo_")


@pytest.fixture(scope="session")
def oracle_api():
    return _oracle_api()


@pytest.fixture(scope="session")
def device_api():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.default_api()


@pytest.fixture(params=["oracle", pytest.param("device", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "oracle":
        return _oracle_api()
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.default_api()
