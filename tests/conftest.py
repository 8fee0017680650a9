import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
ORACLE_EXE = os.path.join(ROOT, "oracle", "_build", "oracle_run_md_simulation")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle is test infrastructure: built here (or prebuilt by __graft_entry__.build())."""
    if not (os.path.exists(ORACLE_LIB) and os.path.exists(ORACLE_EXE)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    return ORACLE_LIB


@pytest.fixture(scope="session")
def cuda_lib():
    from pfmds_b200.build import build
    return build()
