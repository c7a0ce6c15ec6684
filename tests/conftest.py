import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # the oracle (test infrastructure) and the product libraries are built in-tree on demand
    if not os.path.exists(os.path.join(ROOT, "oracle", "build", "libdforacle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    libdir = os.path.join(ROOT, "deep-fusion_b200", "lib")
    if not (os.path.exists(os.path.join(libdir, "libdfcuda.so")) and os.path.exists(os.path.join(libdir, "libdeepfusion.so"))):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "deep-fusion_b200")])


@pytest.fixture(scope="session")
def root():
    return ROOT
