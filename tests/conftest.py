import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def models(oracle):
    from quadruped_locomotion_b200 import legmodel
    return {name: oracle.model_array(legmodel.load_model(name)) for name in ("quadruped_model", "simpledog")}


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_states.npz")))


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat_survey.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def qlb_built():
    """libqlb.so built in-tree (nvcc cross-compiles without a GPU)."""
    from quadruped_locomotion_b200 import build
    return build.build()


def rel_err(a, b):
    """max over components of |a-b| / max(1, |b|_inf) per instance (SURVEY 8d metric)."""
    sc = np.maximum(1.0, np.abs(b).max(axis=0))
    return np.abs(a - b).max(axis=0) / sc
