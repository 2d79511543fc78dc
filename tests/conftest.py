import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker: oracle/liboracle.so (+ oracle/_ref when the reference could be compiled)."""
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the product kernels' source (tests/emu) -- logic check without a GPU."""
    import ctypes
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return ctypes.CDLL(os.path.join(d, "libovvc_b200_emu.so"))


@pytest.fixture(scope="session")
def gpu_ctx():
    """Product path: libovvc_b200.so on cuda:0. Fails loudly when the extension or the GPU is missing."""
    from openvvc_b200 import api
    ctx = api.Ctx(0)
    yield ctx
    ctx.close()
