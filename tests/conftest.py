import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pushing_scene():
    from d3il_b200.scene.blob import load_scene
    return load_scene("pushing")


@pytest.fixture(scope="session")
def avoiding_scene():
    from d3il_b200.scene.blob import load_scene
    return load_scene("avoiding")


@pytest.fixture(scope="session")
def pushing_contexts():
    import numpy as np
    return np.load(os.path.join(ROOT, "d3il_b200", "data", "pushing_test_contexts.npy"))
