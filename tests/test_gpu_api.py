"""The host mirror of the reference interface end to end on the GPU (metacell_b200/api.py): the vignette-a
sequence mat + gset -> mcell_add_cgraph_from_mat_bknn(dsamp = T) -> mcell_coclust_from_graph_resamp, plus the
stand-alone verbs scm_downsamp, gset_get_feat_mat, tgs_cor_knn."""
import numpy as np
import pytest

from conftest import feat_map_of

pytestmark = pytest.mark.gpu


@pytest.fixture()
def db(tmp_path):
    from metacell_b200 import api, synth
    api.scdb_init(str(tmp_path), force_reinit=True)
    m = synth.make_umi_matrix(1300, 500, n_types=6, median_umis=2200.0, seed=31)
    genes = [f"g{i}" for i in range(m.ngenes)]
    cells = [f"cell{i}" for i in range(m.ncells)]
    api.scdb_add_mat("pbmc", api.tgScMat(m.indptr, m.indices, m.data.astype(np.float64), genes=genes, cells=cells))
    feat = [f"g{i}" for i in np.random.default_rng(2).permutation(500)[:180]] + ["not_in_matrix"]
    api.scdb_add_gset("feats", api.tgGeneSets({g: 1 for g in feat}))
    return api, m, feat[:-1]


def test_vignette_sequence(db, ctx, oracle):
    api, m, feat = db
    K = 30
    cg = api.mcell_add_cgraph_from_mat_bknn("pbmc", "feats", "graph", K, dsamp=True, ctx=ctx)
    assert api.scdb_cgraph("graph") is cg and cg.cell_names[0] == "cell0"
    # oracle chain with the same decisions: depth rule, Philox down-sampling with mc_rseed, add back, gene-set order
    tot = m.col_sums()
    n = oracle.downsamp_depth(tot)
    ox, okept = oracle.downsample(m.indptr, m.data, n, api.get_param("mc_rseed"))
    x = np.where(np.repeat(okept, np.diff(m.indptr)), ox, m.data).astype(np.uint32)
    rows = np.array([int(g[1:]) for g in feat], dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, x, feat_map_of(m.ngenes, rows), rows.size)
    col, _ = oracle.cor_knn(Z[valid], K * 10)
    deg, dst, _ = oracle.balance(col, K, 10, 3)
    osrc, odst, ow = oracle.edges_from_balance(deg, dst, K)
    cells = np.nonzero(valid)[0]
    got = set(zip(cg.edges["mc1"].tolist(), cg.edges["mc2"].tolist()))
    exp = set(zip(cells[osrc].tolist(), cells[odst].tolist()))
    assert len(got & exp) / len(got | exp) > 0.999
    assert round(len(cg.edges["w"]) / m.ncells) in (K - 1, K)                 # K = round(nrow(edges)/ncells), R/coclust_graph_cov.r:39
    coc = api.mcell_coclust_from_graph_resamp("coc", "graph", 10, 0.75, 12, ctx=ctx)
    assert coc.n_samp.sum() == 12 * round(0.75 * m.ncells)
    assert np.all(coc.coclust["node1"] < coc.coclust["node2"]) and coc.coclust["cnt"].max() <= 12


def test_scm_downsamp_and_feat_mat(db, ctx, oracle):
    api, m, feat = db
    mat = api.scdb_mat("pbmc")
    n = api.scm_which_downsamp_n(mat, ctx=ctx)
    assert n == oracle.downsamp_depth(m.col_sums())
    ds = api.scm_downsamp(mat, n, ctx=ctx)
    kept = m.col_sums() >= n
    assert ds.ncells == kept.sum() and ds.cells == [c for c, k in zip(mat.cells, kept) if k]
    cs = np.add.reduceat(np.concatenate([ds.data, [0]]), ds.indptr[:-1])[: ds.ncells]
    assert np.all(cs == np.floor(n))                                          # R/utils.r:36-39
    fm = api.gset_get_feat_mat("feats", "pbmc", downsamp=True, add_non_dsamp=True, ctx=ctx)
    assert fm.ngenes == len(feat) and fm.genes == feat and fm.ncells == m.ncells     # gene-set order, all cells back
    with pytest.raises(api.McbError, match="MC-ERR: non existing gset"):
        api.gset_get_feat_mat("nope", "pbmc", ctx=ctx)


def test_tgs_cor_knn_schema(db, ctx, oracle):
    api, m, feat = db
    rows = np.array([int(g[1:]) for g in feat], dtype=np.int32)
    x = np.log2(1.0 + 7.0 * m.to_dense()[rows].astype(np.float64))[:, :400]
    r = api.tgs_cor_knn(x, x, 25, ctx=ctx)
    assert set(r) == {"col1", "col2", "cor", "rank"} and r["col1"].size == 400 * 25        # R/atlas.r:282
    first = r["rank"] == 1
    assert np.array_equal(r["col1"][first], r["col2"][first]) and np.allclose(r["cor"][first], 1.0)   # self at rank 1, R/atlas.r:475-476
    ref = np.corrcoef(x.T)
    assert np.abs(r["cor"] - ref[r["col1"], r["col2"]]).max() < 1e-5
    # y != x (atlas projection, R/atlas.r:164-166): the 5 best of the 10 y cells per x cell, no forced self
    r2 = api.tgs_cor_knn(x, x[:, :10].copy(), 5, ctx=ctx)
    assert r2["col1"].size == 400 * 5 and r2["col2"].max() < 10
    assert np.abs(r2["cor"] - ref[r2["col1"], r2["col2"]]).max() < 1e-5
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgs_cor_knn(x, x[:50, :10].copy(), 5, ctx=ctx)                 # different number of genes
