import os
import sys

import pytest

# tests/test_group_gpu.py puts up to 8 members of an mr_group on ONE device: each member's stream carries a kernel that
# waits for the others, so every stream needs its own hardware queue (the default of 8 connections is shared with
# the other streams of the session).  Must be set before CUDA initialises; irrelevant with one member per GPU.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
