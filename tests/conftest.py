import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/libloops_oracle.so), built on demand."""
    from helpers import Oracle
    so = os.path.join(ROOT, "oracle", "libloops_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle(so)


@pytest.fixture(scope="session")
def ref_host():
    """Unmodified reference headers compiled host-only (oracle/_ref); optional."""
    import ctypes
    so = os.path.join(ROOT, "oracle", "_ref", "libloopsref_host.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libloopsref_host.so not built (reference not mounted)")
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def ref_gpu():
    import ctypes
    so = os.path.join(ROOT, "oracle", "_ref", "libloopsref_gpu.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libloopsref_gpu.so not built (reference not mounted)")
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def battery():
    from helpers import load_battery
    return load_battery()
