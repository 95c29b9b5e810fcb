import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so), built on demand with gcc."""
    from oracle_lib import Oracle

    so = ROOT / "oracle" / "liboracle.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "oracle"], check=True, capture_output=True)
    return Oracle(so)


@pytest.fixture(scope="session")
def ensure_built():
    """The product libraries; built in-tree when missing (nvcc cross-compiles without a GPU)."""
    import montecarlocuda_b200 as m

    if not m.library_path().exists():
        from montecarlocuda_b200 import build as b

        b.build()
    return m


@pytest.fixture(scope="session")
def engine(ensure_built):
    """A pricing context on cuda:0.  No skip, no fallback: without the CUDA library or a device the
    GPU tests fail."""
    m = ensure_built
    eng = m.Engine(int(os.environ.get("MCB200_TEST_DEVICE", "0")))
    yield eng
    eng.close()
