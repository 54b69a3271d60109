"""GPU parity for the SURVEY 8(f) "next" rows that reuse the hot-path kernels: tgs_cor_knn with y != x (atlas
projection), the raw K-nn graph and the metacell confusion matrix."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cells(n, G, seed, n_types=5):
    rng = np.random.default_rng(seed)
    centers = rng.gamma(2.0, 1.0, size=(n_types, G))
    t = rng.integers(0, n_types, n)
    return np.log2(1.0 + 7.0 * rng.poisson(centers[t] * rng.lognormal(0, 0.3, (n, 1)))).astype(np.float64)


@pytest.mark.parametrize("nx,ny,G,knn", [(300, 500, 200, 10), (1000, 77, 64, 50), (130, 4500, 333, 100), (5, 3, 40, 8)])
def test_cor_knn_xy_matches_oracle(ctx, nx, ny, G, knn):
    from metacell_b200 import _lib
    from oracle import oracle_np
    X, Y = _cells(nx, G, 1), _cells(ny, G, 2)
    X[3] = 2.0                                  # constant x cell: NA row
    if ny > 10:
        Y[7] = 0.0                              # constant y cell: never listed
    col, cor = _lib.cor_knn_xy(ctx, X, Y, knn)
    ocol, ocor = oracle_np.cor_knn_xy(X, Y, knn + 4)
    kk = min(knn, ny - (1 if ny > 10 else 0))
    assert np.all(col[3] == -1) and np.all(np.isnan(cor[3]))
    assert np.all(col[:, kk:] == -1)
    rows = np.array([r for r in range(nx) if r != 3])
    # correlations within 1e-5 of the fp64 oracle rank by rank (north_star tolerance)
    assert np.nanmax(np.abs(cor[rows, :kk] - ocor[rows, :kk])) < 1e-5
    # neighbour identity: exact except where the oracle's correlations tie within 2e-5 across the boundary
    if ny > 10:
        assert not np.any(col == 7)
    Zx, _ = oracle_np.standardize_cells(X)
    Zy, _ = oracle_np.standardize_cells(Y)
    for r in rows[:: max(1, rows.size // 60)]:
        got, want = col[r, :kk], ocol[r, :kk]
        if np.array_equal(got, want):
            continue
        exact = Zx[r] @ Zy.T
        assert np.all(np.abs(np.sort(exact[got])[::-1] - np.sort(exact[want])[::-1]) < 2e-5), r
        assert np.all(np.diff(exact[got]) < 2e-5), r                     # still descending up to the tolerance


def test_tgs_cor_knn_xy_mirror_schema(ctx):
    """tgs_cor_knn(x, y, knn) through the R-named mirror: col1/col2/cor/rank (R/atlas.r:282), genes x cells inputs"""
    from metacell_b200 import api
    X, Y = _cells(40, 50, 3), _cells(60, 50, 4)
    r = api.tgs_cor_knn(X.T, Y.T, 5, ctx=ctx)
    assert set(r) == {"col1", "col2", "cor", "rank"}
    assert r["col1"].size == 40 * 5 and r["rank"].min() == 1 and r["rank"].max() == 5
    assert np.array_equal(r["col1"], np.repeat(np.arange(40), 5))
    assert r["col2"].max() < 60
    c = r["cor"].reshape(40, 5)
    assert np.all(np.diff(c, axis=1) <= 0)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgs_cor_knn(X.T, Y[:, :10].T, 5, ctx=ctx)


def test_raw_knn_graph_and_confusion(ctx, oracle, tmp_path):
    """mcell_add_cgraph_from_mat_raw_knn (R/cgraph_knn.r:35-51) against the oracle's kNN, then
    mcell_mc_confusion_mat (R/mc_confusion.r:10-45) against a literal loop"""
    from metacell_b200 import api, synth
    from conftest import feat_map_of
    m = synth.make_umi_matrix(700, 300, n_types=5, median_umis=2500.0, seed=21)
    api.scdb_init(str(tmp_path), force_reinit=True)
    genes = [f"g{i}" for i in range(m.ngenes)]
    cells = [f"c{i}" for i in range(m.ncells)]
    api.scdb_add_mat("m", api.tgScMat(m.indptr, m.indices, m.data, genes=genes, cells=cells, ngenes=m.ngenes))
    feats = np.arange(0, 300, 3, dtype=np.int32)
    api.scdb_add_gset("gs", api.tgGeneSets({genes[g]: 1 for g in feats}))
    K = 12
    gr = api.mcell_add_cgraph_from_mat_raw_knn("m", "gs", "graw", K, ctx=ctx)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    assert valid.all()
    ocol, ocor = oracle.cor_knn(Z, K + 4)
    src, dst, w = gr.edges["mc1"], gr.edges["mc2"], gr.edges["w"]
    assert src.size == m.ncells * K
    assert np.array_equal(src, np.repeat(np.arange(m.ncells), K))
    assert np.allclose(w.reshape(m.ncells, K), 1.0 - np.arange(K) / K)
    got = dst.reshape(m.ncells, K)
    assert np.array_equal(got[:, 0], np.arange(m.ncells))                 # self at rank 1
    exact = Z @ Z.T
    for r in range(m.ncells):
        if not np.array_equal(got[r], ocol[r, :K]):
            assert np.all(np.abs(exact[r, got[r, 1:]] - ocor[r, 1:K]) < 2e-5), r
    # confusion matrix on an arbitrary 4-metacell cover with a few outliers
    rng = np.random.default_rng(5)
    mcid = rng.integers(1, 5, m.ncells)
    mc = {cells[i]: int(mcid[i]) for i in range(m.ncells) if i % 17 != 0}
    api._scdb_add("mc", "mc1", {"mc": mc, "outliers": [cells[i] for i in range(0, m.ncells, 17)]})
    confu = api.mcell_mc_confusion_mat("mc1", "graw", K)
    want = np.zeros((4, 4), dtype=np.int64)
    for a, b in zip(src, dst):
        if cells[a] in mc and cells[b] in mc:
            want[mc[cells[a]] - 1, mc[cells[b]] - 1] += 1
    assert np.array_equal(confu, want)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mcell_mc_confusion_mat("nope", "graw", K)
