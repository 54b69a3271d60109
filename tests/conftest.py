import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    from metacell_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def feat_map_of(ngenes, feats):
    fm = np.full(ngenes, -1, np.int32)
    fm[np.asarray(feats)] = np.arange(len(feats), dtype=np.int32)
    return fm
