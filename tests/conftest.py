import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the native artefacts once if they are missing (they are git-ignored)."""
    lib = os.path.join(ROOT, "physicsbasedanimationtoolkit_b200", "libvbdx.so")
    port = os.path.join(ROOT, "oracle", "liboracle_port.so")
    if not (os.path.exists(lib) and os.path.exists(port)):
        import __graft_entry__ as g

        g.build()
