import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scenes"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    os.chdir(ROOT)   # scene files name their assets relative to the repository root


@pytest.fixture(scope="session", autouse=True)
def assets():
    import gen_assets
    return gen_assets.ensure_assets()


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (they are built by __graft_entry__.build())."""
    lib = os.path.join(ROOT, "bhraytracer_b200", "lib")
    need = [os.path.join(lib, "libbhrt.so"), os.path.join(lib, "libbhrt_host.so"), os.path.join(ROOT, "build", "emu_render"), os.path.join(ROOT, "build", "desc_dump")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "build"])
    return lib


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")
EMU_BIN = os.path.join(ROOT, "build", "emu_render")


def have_ref():
    return os.path.exists(REF_BIN)
