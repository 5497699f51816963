import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import subprocess
    import _b200pkg
    lib_path = os.path.join(ROOT, "minisat-rust_b200", "libb200sat.so")
    if not os.path.exists(lib_path):   # fresh checkout: the built .so files are git-ignored
        subprocess.run(["make", "-C", os.path.join(ROOT, "minisat-rust_b200", "csrc")], check=True)
    return _b200pkg.load()


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
