import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cport():
    """The plain-C oracle (oracle/pt_oracle.c); built on demand, gcc is everywhere."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.CPORT_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), pyoracle.CPORT_PATH])
    return pyoracle.CPort()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in place (oracle/_ref/libptref.so), when it has been built."""
    from oracle import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libptref.so not built (needs /root/reference at build time)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def golden_renders():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_renders.npz"))


@pytest.fixture(scope="session")
def c1():
    import scenes
    return scenes.load_c1()
