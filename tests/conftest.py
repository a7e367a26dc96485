import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p_ in (ROOT, HERE):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def native_build():
    """Build the CPU-side native pieces once per session (seconds): world producer, C oracle, host emulator, and the
    verbatim-reference oracle when /root/reference is present.  libgcmesh.so is built by __graft_entry__.build()."""
    from gamecraft_b200 import worldgen
    from oracle import binding as ob

    worldgen.build()
    ob.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    yield
