import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One mr_ctx on cuda:0 for the whole GPU test session (fails loudly without a GPU)."""
    import metarank_b200 as mb

    c = mb.Context(0)
    yield c
    c.close()
