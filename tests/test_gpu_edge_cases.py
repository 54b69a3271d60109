"""Edge cases of the path on the GPU: tiny / empty / ragged inputs, degenerate cells, maximum-ish list sizes."""
import numpy as np
import pytest

from conftest import feat_map_of

pytestmark = pytest.mark.gpu


def _mat(ncells, ngenes=200, seed=9, **kw):
    from metacell_b200 import synth
    return synth.make_umi_matrix(ncells, ngenes, n_types=4, median_umis=800.0, seed=seed, **kw)


def _graph_via_gpu(ctx, m, feats, K, kx=10, kb=3, data=None):
    from metacell_b200 import _lib
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, (m.data if data is None else data).astype(np.float64))
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(K, kx)
    col, cor = g.download_knn()
    g.mutual(kb)
    g.prune()
    out = (col, cor, g.edges(), g.valid_cells())
    g.destroy()
    return out


@pytest.mark.parametrize("ncells,K,kx", [(2, 1, 1), (3, 5, 10), (17, 4, 10), (130, 7, 3), (257, 100, 10)])
def test_tiny_and_ragged_sizes(ctx, oracle, ncells, K, kx):
    m = _mat(ncells)
    feats = np.arange(0, 200, 3, dtype=np.int32)                    # G = 67: not a multiple of the 64-wide k-block
    col, cor, (src, dst, w), cells = _graph_via_gpu(ctx, m, feats, K, kx)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    assert np.array_equal(cells, np.nonzero(valid)[0])
    ocol, ocor = oracle.cor_knn(Z[valid], K * kx)
    assert col.shape == ocol.shape
    assert np.abs(cor - ocor).max() < 1e-5
    same = (col == ocol).mean()
    assert same > 0.99, same
    deg, odst, _ = oracle.balance(col, K, kx, 3)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    assert np.array_equal(src, cells[osrc]) and np.array_equal(dst, cells[odst2]) and np.array_equal(w, ow)


def test_two_feature_genes(ctx, oracle):
    m = _mat(300)
    feats = np.array([5, 9], dtype=np.int32)                         # G = 2: correlations are +-1, massive ties
    col, cor, (src, dst, w), cells = _graph_via_gpu(ctx, m, feats, 5, 4)
    assert np.all(np.abs(np.abs(cor[col >= 0]) - 1.0) < 1e-5)
    assert np.all(col[:, 0] == np.arange(cells.size))


def test_all_cells_degenerate_is_an_error(ctx):
    from metacell_b200 import _lib
    m = _mat(50)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, np.zeros(m.nnz))
    g = _lib.CGraph.from_csc(ctx, csc, np.arange(20, dtype=np.int32), 7.0)
    assert g.M == 0
    with pytest.raises(_lib.McbError, match="MC-ERR"):
        g.knn(5, 10)
    g.destroy()


def test_empty_matrix_and_empty_cells(ctx):
    from metacell_b200 import _lib
    indptr = np.zeros(11, dtype=np.int64)
    csc = _lib.Csc.upload(ctx, 30, indptr, np.zeros(0, np.int32), np.zeros(0))
    assert np.array_equal(csc.col_sums(), np.zeros(10, np.int64))
    ds = csc.downsample(5.0, 1)
    x, kept = ds.download()
    assert x.size == 0 and not kept.any()
    g = _lib.CGraph.from_csc(ctx, csc, np.arange(10, dtype=np.int32), 7.0)
    assert g.M == 0
    g.destroy()


def test_downsample_depth_above_every_cell(ctx, oracle):
    from metacell_b200 import _lib
    m = _mat(64)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    n = float(m.col_sums().max() + 1)
    x, kept = csc.downsample(n, 3).download()
    assert not kept.any() and not x.any()
    x2, kept2 = csc.downsample(n, 3, add_non_dsamp=True).download()
    assert not kept2.any() and np.array_equal(x2, m.data.astype(np.float64))     # raw counts re-enter (R/gset.r:179-187)


def test_large_counts_downsample(ctx, oracle):
    """a very deep cell (1e6 UMIs in 3 genes) next to shallow ones: bit-exact with the oracle"""
    from metacell_b200 import _lib
    indptr = np.array([0, 3, 5, 6], dtype=np.int64)
    indices = np.array([0, 1, 2, 0, 2, 1], dtype=np.int32)
    data = np.array([600000, 300000, 100000, 40, 2, 7], dtype=np.uint32)
    csc = _lib.Csc.upload(ctx, 3, indptr, indices, data.astype(np.float64))
    for n in (7.0, 42.0, 1000.0):
        x, kept = csc.downsample(n, 11).download()
        ox, okept = oracle.downsample(indptr, data, n, 11)
        assert np.array_equal(kept, okept) and np.array_equal(x.astype(np.uint32), ox)


def test_coclust_edge_cases(ctx, oracle):
    from metacell_b200 import _lib, synth
    N = 300
    sets, seeds = synth.resample_sets(N, 1.0, 3, 5)                   # p = 1: every node sampled
    # empty graph: nobody covered, no pairs
    cc = _lib.CoClust(ctx, N, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), sets, seeds, 5, 1.05, 10)
    n1, n2, cnt, ns = cc.pairs()
    assert n1.size == 0 and np.all(ns == 3)
    mc, sw = cc.assignments()
    assert np.all(mc == -1)
    cc.destroy()
    # min_mc_size larger than any neighbourhood: no seed can be drawn -> all outliers, bit-exact with the oracle
    m = _mat(N)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, np.arange(200, dtype=np.int32), 200)
    col, _ = oracle.cor_knn(Z[valid], 40)
    deg, dst, _ = oracle.balance(col, 8, 5, 3)
    src, d, w = oracle.edges_from_balance(deg, dst, 8)
    M = int(valid.sum())
    sets, seeds = synth.resample_sets(M, 0.6, 4, 9)
    for min_size in (1, 6, 50):
        cc = _lib.CoClust(ctx, M, src, d, w, sets, seeds, min_size, 1.05, 10)
        n1, n2, cnt, ns = cc.pairs()
        csr = oracle.graph_to_csr(M, src, d, w)
        ocnt, ons, _ = oracle.coclust(M, csr, sets, seeds, min_size, 1.05, 10)
        dense = np.zeros((M, M), np.uint32)
        dense[n1, n2] = cnt
        assert np.array_equal(dense, ocnt) and np.array_equal(ns, ons), min_size
        cc.destroy()
    # unsorted sample sets are rejected (contract of mcb_coclust_run)
    bad = sets.copy()
    bad[0, :2] = bad[0, 1::-1]
    with pytest.raises(_lib.McbError, match="MC-ERR"):
        _lib.CoClust(ctx, M, src, d, w, bad, seeds, 5, 1.05, 10)
