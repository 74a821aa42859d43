import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: full-length golden runs (FQSB_FULL_GOLDEN=1)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


def pytest_sessionstart(session):
    """A clean checkout has no built artefacts (they are git-ignored): build them once."""
    lib = ROOT / "frictionqpotspringblock_b200" / "libfqsb.so"
    ora = ROOT / "oracle" / "libfqsb_oracle.so"
    if not lib.exists() or not ora.exists():
        import __graft_entry__

        __graft_entry__.build()
