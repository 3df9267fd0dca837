"""pytest configuration: markers and shared paths.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol checks (no GPU needed).
``-m gpu``      : parity tests proper; they call the CUDA path through the C-ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
