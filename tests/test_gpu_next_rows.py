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


def _dense(m, data=None):
    d = np.zeros((m.ngenes, m.ncells))
    x = m.data if data is None else data
    for c in range(m.ncells):
        s, e = m.indptr[c], m.indptr[c + 1]
        d[m.indices[s:e], c] = x[s:e]
    return d


@pytest.mark.parametrize("ncells,ngenes,q,use_sub", [(400, 300, 0.2, False), (777, 150, 0.1, True), (64, 40, 0.5, False)])
def test_gene_stats_match_dense_restatement(ctx, ncells, ngenes, q, use_sub):
    """mcb_gene_stats against the literal dense restatement of R/gstats.r:83-151: integer statistics exact, floating
    point ones to 1e-10 relative (the reference rounds them to 3-7 decimals)"""
    from metacell_b200 import _lib, synth
    from oracle import oracle_np
    m = synth.make_umi_matrix(ncells, ngenes, n_types=4, median_umis=900.0, seed=ncells)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    csum = csc.col_sums()
    n = float(np.quantile(csum, 0.3))                               # fractional depth, drops ~30 % of the cells
    ds = csc.downsample(n, 19, add_non_dsamp=False)
    dsx, kept = ds.download()
    sub = np.sort(np.random.default_rng(3).choice(ncells, ncells // 2, replace=False)).astype(np.int32) if use_sub else None
    got = _lib.gene_stats(ctx, csc, ds, q, 1000.0, sub)
    want = oracle_np.gene_stats_dense(_dense(m), _dense(m, dsx)[:, kept], q, 1000.0, sub)
    for k in ("tot", "is_on_count", "ds_top1", "ds_top2", "ds_top3", "ds_is_on_count"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("var", "niche_stat", "n_mean", "ds_mean", "ds_var", "sz_cor"):
        a, b = got[k], want[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        ok = ~np.isnan(a)
        assert np.allclose(a[ok], b[ok], rtol=1e-10, atol=1e-12), (k, np.abs(a[ok] - b[ok]).max())
    # without a downsampled matrix the ds_* columns are NA
    got2 = _lib.gene_stats(ctx, csc, None, q, 1000.0, sub)
    assert np.all(np.isnan(got2["ds_mean"])) and np.array_equal(got2["tot"], got["tot"])
    ds.free(); csc.free()


def test_scm_gene_stat_mirror(ctx, tmp_path):
    """scm_gene_stat through the R-named mirror: schema (R/gstats.r:195-199), gene filter, rounding and the
    rolling-median trends against oracle_np.rollmedian_fill"""
    from metacell_b200 import api, synth
    from oracle import oracle_np
    m = synth.make_umi_matrix(1500, 400, n_types=6, median_umis=1500.0, seed=8)
    api.scdb_init(str(tmp_path), force_reinit=True)
    api.scdb_add_mat("m", api.tgScMat(m.indptr, m.indices, m.data, ngenes=m.ngenes))
    gs = api.mcell_add_gene_stat("m", "gs", ctx=ctx)
    cols = ["name", "tot", "var", "is_on_count", "sz_cor", "sz_cor_norm", "niche_stat", "niche_norm", "n_mean", "ds_top1",
            "ds_top2", "ds_top3", "ds_mean", "ds_var", "ds_log_varmean", "ds_vm_norm", "ds_is_on_count", "downsample_n"]
    assert list(gs.keys()) == cols
    ng = len(gs["name"])
    assert all(len(gs[k]) == ng for k in cols) and ng >= 102
    assert np.all(gs["tot"] > 10)
    assert np.all(gs["ds_top1"] >= gs["ds_top2"]) and np.all(gs["ds_top2"] >= gs["ds_top3"])
    assert np.all(gs["niche_stat"] > 0) and np.all(gs["niche_stat"] <= 1)
    assert np.allclose(gs["var"], np.round(gs["var"], 7)) and np.allclose(gs["sz_cor"], np.round(gs["sz_cor"], 3))
    order = np.argsort(gs["tot"], kind="stable")
    v = gs["niche_stat"][order]
    trend = oracle_np.rollmedian_fill(v, 101, np.median(v[:101]), np.median(v[-102:]))
    assert np.allclose(gs["niche_norm"][order], v - trend, atol=1e-12)
    assert api._scdb_get("gstat", "gs") is gs
    with pytest.raises(api.McbError):
        api.scm_gene_stat("m", niche_quantile=1.5, ctx=ctx)


def test_mc_footprints_match_dense_restatement(ctx, tmp_path):
    """mc_compute_fp / e_gc / cov_gc (R/mc.r:164-241) through the mirror against the literal dense restatement;
    a few cells are outliers (outside the cover) and one count is large enough to leave the kernel's lookup table"""
    from metacell_b200 import api, synth
    from oracle import oracle_np
    m = synth.make_umi_matrix(900, 250, n_types=5, median_umis=1200.0, seed=31)
    data = m.data.copy()
    data[5] = 5000.0
    rng = np.random.default_rng(2)
    ids = rng.integers(0, 7, m.ncells)
    ids[::13] = -1
    cells = [f"c{i}" for i in range(m.ncells)]
    mat = api.tgScMat(m.indptr, m.indices, data, cells=cells, ngenes=m.ngenes)
    mc = {"mc": {cells[i]: int(ids[i]) + 1 for i in range(m.ncells) if ids[i] >= 0}, "outliers": [cells[i] for i in range(0, m.ncells, 13)]}
    dense = _dense(m, data)
    cov = ids >= 0
    geo, frac, msz = oracle_np.mc_tapply_dense(dense[:, cov], ids[cov])
    f_g = dense[:, cov].sum(axis=1) > 10
    e, genes = api.mc_compute_e_gc(mc, mat, ctx=ctx)
    assert genes == [mat.genes[i] for i in np.nonzero(f_g)[0]]
    assert np.allclose(e, geo[f_g] / msz[None, :], rtol=1e-11, atol=1e-14)
    c, _ = api.mc_compute_cov_gc(mc, mat, ctx=ctx)
    assert np.array_equal(c, frac[f_g])
    fp, _ = api.mc_compute_fp(mc, mat, ctx=ctx)
    g = min(1000.0, np.median(msz)) * geo[f_g] / msz[None, :]
    want = (0.1 + g) / np.median(0.1 + g, axis=1, keepdims=True)
    assert np.allclose(fp, want, rtol=1e-10)
    upd = api.mc_update_stats(mc, mat, ctx=ctx)
    assert np.array_equal(upd["mc_fp"], fp) and np.array_equal(upd["cov_gc"], c)      # deterministic: fixed-point sums
    bad = dict(mc); bad["mc"] = dict(mc["mc"]); bad["mc"]["nope"] = 1
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mc_compute_fp(bad, mat, ctx=ctx)
