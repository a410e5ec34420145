import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emul_lib():
    """Host emulation of the kernel bodies (tests/host/emul.cpp), built on demand with g++."""
    import ctypes
    src = os.path.join(ROOT, "tests", "host", "emul.cpp")
    out = os.path.join(ROOT, "tests", "host", "libplb_emul.so")
    deps = [src] + [os.path.join(ROOT, "plasticinelab_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "plasticinelab_b200", "csrc"))]
    if not os.path.isfile(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", out])
    return ctypes.CDLL(out)
