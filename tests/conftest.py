"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tiny-ntt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
TAGS = ("dilithium", "n1024_24", "n4096_24", "n4096_60")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(tag):
    with gzip.open(os.path.join(GOLDEN_DIR, f"golden_{tag}.json.gz"), "rt") as fh:
        return json.load(fh)


@pytest.fixture(scope="session", params=TAGS)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope="session")
def golden_all():
    return {t: load_golden(t) for t in TAGS}
