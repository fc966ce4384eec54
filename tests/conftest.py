import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_fe():
    return np.load(os.path.join(GOLDEN, "frontend_golden.npz"))


@pytest.fixture(scope="session")
def golden_knn():
    return np.load(os.path.join(GOLDEN, "knn_golden.npz"))


@pytest.fixture(scope="session")
def golden_seq():
    """Sequence-feature KNN of compare_feature_methods.py, captured from the reference (oracle/gen_golden_seq.py)."""
    return np.load(os.path.join(GOLDEN, "knn_seq_golden.npz"))


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library; fails loudly (no CPU fallback) without a GPU."""
    from dsp_audioreclabs_b200 import batch
    return batch.default_context(0)


def golden_pcm(g, name):
    from oracle import synth
    if name.startswith("full"):
        return synth.utterance_pcm(int(name[4:]))
    return g[f"pcm/{name}"]


def golden_names(g):
    return list(g["names"]) + [f"full{i}" for i in g["full_index"]]
