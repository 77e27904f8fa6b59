import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:          # test modules share helpers (from test_gpu_parity import sha, check_eigs)
    sys.path.insert(0, HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_tiny():
    return dict(np.load(os.path.join(GOLDEN, "tiny.npz")))


@pytest.fixture(scope="session")
def golden_tiny_masked():
    return dict(np.load(os.path.join(GOLDEN, "tiny_masked.npz")))


@pytest.fixture(scope="session")
def golden_c1():
    return dict(np.load(os.path.join(GOLDEN, "c1.npz")))


@pytest.fixture(scope="session")
def golden_tiny_stageF():
    return dict(np.load(os.path.join(GOLDEN, "tiny_stageF.npz")))


@pytest.fixture(scope="session")
def golden_c1_stageF():
    return dict(np.load(os.path.join(GOLDEN, "c1_stageF.npz")))


@pytest.fixture(scope="session")
def golden_tiny_stageG():
    return dict(np.load(os.path.join(GOLDEN, "tiny_stageG.npz")))


@pytest.fixture(scope="session")
def golden_c1_stageG():
    return dict(np.load(os.path.join(GOLDEN, "c1_stageG.npz")))


@pytest.fixture(scope="session")
def golden_tiny_stageH():
    return dict(np.load(os.path.join(GOLDEN, "tiny_stageH.npz")))


@pytest.fixture(scope="session")
def golden_c1_stageH():
    return dict(np.load(os.path.join(GOLDEN, "c1_stageH.npz")))


@pytest.fixture(scope="session")
def corpus_c1():
    from isle_b200 import corpus
    return corpus.generate("c1")


@pytest.fixture(scope="session")
def ctx():
    """One device context for the GPU tests; fails loudly when the extension or GPU is missing."""
    from isle_b200 import _capi
    c = _capi.Context(0)
    yield c
    c.close()
