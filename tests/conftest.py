import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dkb():
    """The package with libdkb.so built (CPU-only build works: nvcc cross-compiles)."""
    from denovo_kmer_b200 import build
    build.build()
    import denovo_kmer_b200
    return denovo_kmer_b200


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle
