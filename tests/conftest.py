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


def knn_parity(col, cor, ocol_x, ocor_x, tol=1e-5):
    """ocol_x / ocor_x: oracle lists computed a few ranks deeper than `col`, so that the tie partner of
    the last rank is visible.  Disagreements are only allowed inside runs of correlations tied within tol."""
    kc = col.shape[1]
    ocol, ocor = ocol_x[:, :kc], ocor_x[:, :kc]
    assert np.abs(cor - ocor).max() < tol
    assert np.array_equal(col[:, 0], ocol[:, 0])                      # self is rank 1 (R/atlas.r:475-476)
    diff = col != ocol
    if diff.any():
        r, q = np.nonzero(diff)
        for i, j in zip(r, q):
            nb = [abs(ocor_x[i, j] - ocor_x[i, jj]) for jj in (j - 1, j + 1) if 0 < jj < ocor_x.shape[1] and ocol_x[i, jj] >= 0]
            assert nb and min(nb) < 2 * tol, (i, j, ocor_x[i, max(j - 1, 0):j + 2], cor[i, max(j - 1, 0):j + 2])
        for i in np.unique(r):
            assert len(set(col[i]) ^ set(ocol[i])) <= 2 * int(diff[i].sum())
    return float((~diff).mean())
