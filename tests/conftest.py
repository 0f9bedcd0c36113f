import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so), built on demand.  Test infrastructure only."""
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def rsf():
    import rsfusion
    return rsfusion


@pytest.fixture(scope="session")
def synth():
    import rsfusion  # noqa: F401  (registers the package)
    import importlib
    return importlib.import_module("ra-slam_b200.synth")


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
