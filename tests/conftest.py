import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running full-size case")


@pytest.fixture(scope="session")
def oracle_bin():
    """Path of the oracle CLI (built on demand; CPU only, test infrastructure)."""
    path = os.path.join(ROOT, "oracle", "_build", "oracle_cornetto")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return path


@pytest.fixture(scope="session")
def ref_bin():
    """The unmodified reference binary compiled by oracle/Makefile, if present."""
    path = os.path.join(ROOT, "oracle", "_ref", "cornetto")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        if not os.path.exists(path):
            pytest.skip("oracle/_ref/cornetto not available")
    return path
