"""
pytest configuration.

`-m "not gpu"` : the CPU oracle against the golden vectors, host logic, C-ABI surface (no GPU).
`-m gpu`       : parity tests proper -- the sm_100a library, called through the C-ABI, against the
                 oracle and the committed golden fixtures.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robotics-rl-srl_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle_sim.so")
CUDA_LIB = os.path.join(PKG, "csrc", "libsrl_sim_b200.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_oracle():
    # building the checker is not using it; idempotent (make is a no-op when up to date)
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


@pytest.fixture(scope="session")
def oracle_lib():
    from srl_sim._abi import SimLibrary
    _build_oracle()
    return SimLibrary(ORACLE_LIB)


@pytest.fixture(scope="session")
def oracle_backend(oracle_lib):
    from srl_sim.backend import Backend
    return Backend(oracle_lib, -1)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from srl_sim._abi import load_cuda_library
    return load_cuda_library()  # raises if the extension is missing: no silent fallback


@pytest.fixture(scope="session")
def cuda_backend(cuda_lib):
    from srl_sim.backend import Backend
    return Backend(cuda_lib, 0)


@pytest.fixture()
def use_oracle_backend(oracle_lib):
    """Route the host-side env classes to the oracle library for the duration of one CPU test."""
    from srl_sim import backend
    backend.use_library(oracle_lib, -1)
    yield
    backend.use_library(None, None)
