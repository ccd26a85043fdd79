"""pytest configuration: the `gpu` marker, repo-root imports and shared fixtures.

`-m "not gpu"` : oracle vs the reference's golden vectors, host logic, host simulation of the device
                 code, C-ABI surface (no compute).  Runs anywhere, a few minutes.
`-m gpu`       : the parity tests proper, through the C ABI of libeccbatch.so on cuda:0.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (loudly), not errored: the library has no CPU
    fallback and nothing here may pretend otherwise."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the hot path has no CPU fallback (run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def coracle():
    from oracle import coracle as C

    C.build()
    C.load()
    return C


@pytest.fixture(scope="session")
def ctx():
    """One libeccbatch context on device 0.  No fallback: a missing library or device is an error."""
    from eccoxide_b200 import Context

    c = Context()
    yield c
    c.close()
