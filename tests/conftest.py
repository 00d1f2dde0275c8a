import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rs():
    import recsys_b200
    recsys_b200._lib.load()
    return recsys_b200


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build_chain_lib()
    return oracle
