"""GPU parity tests of the balanced-graph path (A1..A7) against the CPU oracle, through the C ABI.

Tolerances (BASELINE.json north_star): correlations within 1e-5 absolute; kNN and balanced-graph
edge sets bit-exact except at correlation ties within that tolerance; down-sampled counts bit-exact.
"""
import numpy as np
import pytest

from conftest import feat_map_of, knn_parity as _knn_parity

pytestmark = pytest.mark.gpu

COR_TOL = 1e-5


def _small(ncells=1200, ngenes=500, seed=3, **kw):
    from metacell_b200 import synth
    return synth.make_umi_matrix(ncells, ngenes, n_types=7, median_umis=2000.0, seed=seed, **kw)


def _upload(ctx, m, data=None):
    from metacell_b200 import _lib
    return _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, (m.data if data is None else data).astype(np.float64))


# ---- A1 / A2 -----------------------------------------------------------------------------------
def test_col_sums_and_depth(ctx, oracle):
    from metacell_b200 import _lib
    m = _small()
    csc = _upload(ctx, m)
    tot = csc.col_sums()
    assert np.array_equal(tot, m.col_sums())
    assert _lib.downsamp_depth(tot) == oracle.downsamp_depth(tot)


@pytest.mark.parametrize("n", [700.0, 1500.5, 2600.0])
def test_downsample_bit_exact(ctx, oracle, n):
    m = _small()
    csc = _upload(ctx, m)
    ds = csc.downsample(n, 42, add_non_dsamp=False)
    x, kept = ds.download()
    ox, okept = oracle.downsample(m.indptr, m.data, n, 42)
    assert np.array_equal(kept, okept)
    assert np.array_equal(x.astype(np.uint32), ox)
    # invariants of R/utils.r:34-39
    cs = np.add.reduceat(np.concatenate([x, [0]]), m.indptr[:-1])[: m.ncells]
    cs[np.diff(m.indptr) == 0] = 0
    assert np.all(cs[kept] == np.floor(n))
    assert np.all(x <= m.data)
    assert np.all(cs[~kept] == 0)


def test_downsample_add_back_and_edge_cases(ctx, oracle):
    from metacell_b200 import synth
    m = _small(ncells=300)
    # an empty cell, a cell exactly at depth, a very deep cell
    data = m.data.copy()
    c0 = slice(m.indptr[0], m.indptr[1])
    data[c0] = 0
    n = 1000.0
    c1 = slice(m.indptr[1], m.indptr[2])
    data[c1] = 0
    data[m.indptr[1]] = 1000
    data[m.indptr[2]] = 200000
    csc = _upload(ctx, m, data)
    ds = csc.downsample(n, 7, add_non_dsamp=True)
    x, kept = ds.download()
    ox, okept = oracle.downsample(m.indptr, data, n, 7)
    assert np.array_equal(kept, okept)
    assert not kept[0] and kept[1] and kept[2]
    exp = np.where(np.repeat(okept, np.diff(m.indptr)), ox, data)     # raw counts re-enter (R/gset.r:179-187)
    assert np.array_equal(x.astype(np.uint32), exp.astype(np.uint32))


# ---- A3 / A4 -----------------------------------------------------------------------------------
def test_features_match_oracle(ctx, oracle):
    from metacell_b200 import _lib
    m = _small()
    feats = np.arange(3, 483, 3, dtype=np.int32)         # 160 features -> Gpad = 160
    data = m.data.copy()
    data[m.indptr[5]:m.indptr[6]] = 0                      # degenerate cell: no feature counts
    csc = _upload(ctx, m, data)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    Z, valid = oracle.features(m.indptr, m.indices, data, feat_map_of(m.ngenes, feats), feats.size)
    assert not valid[5]
    assert np.array_equal(g.valid_cells(), np.nonzero(valid)[0])
    z = g.features(feats.size)
    assert np.abs(z - Z[valid]).max() < 2e-7
    g.destroy()


# ---- the contraction itself ----------------------------------------------------------------------
@pytest.mark.parametrize("G", [96, 600, 1000])
def test_correlation_kernels_vs_fp64(ctx, oracle, G):
    """tcgen05 3-pass split-TF32 block == SIMT fp32 block == fp64 within 1e-5 (measured, printed)."""
    from metacell_b200 import _lib
    m = _small(ncells=700, ngenes=1000)
    feats = np.arange(G, dtype=np.int32)
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), G)
    Zv = Z[valid]
    ref = Zv @ Zv.T
    a, b = min(g.M, 300), g.M
    simt = g.debug_cor(1, a, b)
    tc3 = g.debug_cor(0, a, b)
    tc1 = g.debug_cor(2, a, b)
    e_simt = np.abs(simt - ref[:a, :b]).max()
    e_tc3 = np.abs(tc3 - ref[:a, :b]).max()
    e_tc1 = np.abs(tc1 - ref[:a, :b]).max()
    print(f"G={G}: max|err| simt {e_simt:.2e}  tc 3-pass {e_tc3:.2e}  tc 1-pass {e_tc1:.2e}")
    assert e_simt < COR_TOL
    assert e_tc3 < COR_TOL
    assert e_tc1 < 2e-3
    g.destroy()


# ---- A5 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("ncells,K,kx", [(1200, 20, 10), (900, 100, 10), (3000, 30, 10)])
def test_cor_knn_parity(ctx, oracle, impl, ncells, K, kx):
    from metacell_b200 import _lib
    m = _small(ncells=ncells)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m)
    ctx.set_option("gemm_impl", impl)
    try:
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        g.knn(K, kx)
        col, cor = g.download_knn()
        stats = g.knn_stats()
    finally:
        ctx.set_option("gemm_impl", 0)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    ocol, ocor = oracle.cor_knn(Z[valid], K * kx + 8)
    agree = _knn_parity(col, cor, ocol, ocor)
    print(f"impl={impl} N={ncells} Kc={K * kx}: agreement {agree:.6f} stats {stats}")
    g.destroy()


@pytest.mark.parametrize("ncells,K,kx", [(6000, 70, 10), (9000, 150, 10)])
def test_cor_knn_parity_large_lists(ctx, oracle, ncells, K, kx):
    """candidate capacities above 1,024 and 2,048 keys: the histogram-select + counting-sort ranking kernels
    (topk_select_rows_kernel<8,4> and <16,8>) that the C2 / C4 configurations use, and the balancing kernels at
    Kc = 1,500"""
    from metacell_b200 import _lib
    m = _small(ncells=ncells)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(K, kx)
    col, cor = g.download_knn()
    stats = g.knn_stats()
    assert stats["cap"] > (2048 if K * kx > 1024 else 1024), stats
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    ocol, ocor = oracle.cor_knn(Z[valid], K * kx + 8)
    agree = _knn_parity(col, cor, ocol, ocor)
    print(f"N={ncells} Kc={K * kx}: agreement {agree:.6f} stats {stats}")
    # balancing on the GPU's own lists must equal the oracle's balancing of the same lists, bit for bit
    g.mutual(3)
    g.prune()
    src, dst, w = g.edges()
    deg, odst, _ = oracle.balance(col, K, kx, 3)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    cells = g.valid_cells()
    assert np.array_equal(src, cells[osrc]) and np.array_equal(dst, cells[odst2]) and np.array_equal(w, ow)
    g.destroy()


@pytest.mark.parametrize("impl", [0, 1])
def test_cor_knn_exact_row_path(ctx, oracle, impl):
    """rows whose candidate count misses the window are redone exactly: force it for most rows"""
    from metacell_b200 import _lib
    m = _small(ncells=1500)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m)
    ctx.set_option("gemm_impl", impl)
    ctx.set_option("cap_override", 230)
    try:
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        g.knn(20, 10)
        col, cor = g.download_knn()
        stats = g.knn_stats()
    finally:
        ctx.set_option("gemm_impl", 0)
        ctx.set_option("cap_override", 0)
    assert stats["fallback_rows"] > 100, stats
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    ocol, ocor = oracle.cor_knn(Z[valid], 208)
    agree = _knn_parity(col, cor, ocol, ocor)
    print(f"exact-row path impl={impl}: {stats['fallback_rows']} rows, agreement {agree:.6f}")
    g.destroy()


def test_cor_knn_duplicate_cells_and_ties(ctx, oracle):
    """exact duplicate cells give exactly tied correlations: self still rank 1, ties by lower index"""
    from metacell_b200 import _lib
    m = _small(ncells=600)
    # duplicate cell 10 into cell 11 by rebuilding the CSC
    lens = np.diff(m.indptr)
    cols = [np.arange(m.indptr[c], m.indptr[c + 1]) for c in range(m.ncells)]
    cols[11] = cols[10]
    idx = np.concatenate(cols)
    indptr = np.concatenate([[0], np.cumsum([len(c) for c in cols])])
    from metacell_b200 import synth
    m2 = synth.UmiMatrix(indptr=indptr, indices=m.indices[idx], data=m.data[idx], ngenes=m.ngenes, ncells=m.ncells, cell_type=m.cell_type)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m2)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(10, 10)
    col, cor = g.download_knn()
    assert col[10, 0] == 10 and col[11, 0] == 11
    assert col[10, 1] == 11 and col[11, 1] == 10
    assert abs(cor[10, 1] - 1.0) < COR_TOL
    Z, valid = oracle.features(m2.indptr, m2.indices, m2.data, feat_map_of(m2.ngenes, feats), feats.size)
    ocol, ocor = oracle.cor_knn(Z[valid], 108)
    _knn_parity(col, cor, ocol, ocor)
    g.destroy()


@pytest.mark.parametrize("ncells,ngenes,K", [(20000, 300, 20), (33000, 520, 30), (12900, 200, 40)])
def test_fused_sampling_pass_equals_stored_block_path(ctx, ncells, ngenes, K):
    """The sampling pass only steers the filter: the fused e4m3 histogram kernel (corsample.cu) and the stored fp16 block +
    threshold kernel must give identical ranked lists, a similar number of candidates, and (next to) no exact-row fallbacks."""
    from metacell_b200 import _lib, synth
    m = synth.make_umi_matrix(ncells, ngenes, n_types=12, median_umis=1500.0, seed=5)
    feats = np.arange(ngenes, dtype=np.int32)
    csc = _upload(ctx, m)
    res = {}
    try:
        for impl in (1, -1):
            ctx.set_option("sample_impl", impl)
            g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
            g.knn(K, 10)
            res[impl] = (g.download_knn(), g.knn_stats())
            g.destroy()
    finally:
        ctx.set_option("sample_impl", -1)
    (c1, v1), s1 = res[1]
    (c2, v2), s2 = res[-1]
    assert s1["sample_cols"] >= 4096, "shape too small to reach the fused path"
    assert np.array_equal(c1, c2) and np.allclose(v1, v2, rtol=0, atol=1e-6)     # exact-row fallbacks sum in another order
    assert s2["fallback_rows"] <= max(2, ncells // 5000), s2
    assert 0.85 * s1["candidates"] <= s2["candidates"] <= 1.15 * s1["candidates"], (s1, s2)


# ---- A6 / A7 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("ncells,K,kx,kb", [(1200, 20, 10, 3), (900, 100, 10, 3), (2500, 15, 4, 2)])
def test_balance_bit_exact_on_same_lists(ctx, oracle, ncells, K, kx, kb):
    """Integer stage: balancing the GPU's own kNN lists must equal the oracle bit for bit."""
    from metacell_b200 import _lib
    m = _small(ncells=ncells)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(K, kx)
    col, _ = g.download_knn()
    g.mutual(kb)
    g.prune()
    src, dst, w = g.edges()
    cells = g.valid_cells()
    deg, odst, _ = oracle.balance(col, K, kx, kb)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    assert np.array_equal(src, cells[osrc])
    assert np.array_equal(dst, cells[odst2])
    assert np.array_equal(w, ow)
    # invariants: degrees, weights, threshold
    assert np.bincount(src, minlength=ncells).max() <= K
    assert np.bincount(dst, minlength=ncells).max() <= K * kb
    assert w.max() == 1.0 and w.min() > 0.0
    assert np.all(np.isclose((1.0 - w) * K, np.rint((1.0 - w) * K)))
    g.destroy()


def test_full_graph_parity_end_to_end(ctx, oracle):
    """CSC -> downsample -> features -> kNN -> balance on the GPU vs the same chain in the oracle."""
    from metacell_b200 import _lib
    m = _small(ncells=1500, ngenes=800)
    feats = np.sort(np.random.default_rng(0).choice(800, 300, replace=False)).astype(np.int32)
    K, kx, kb = 25, 10, 3
    src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, K, kx, kb,
                                               7.0, dsamp=True, seed=42)
    tot = m.col_sums()
    n = oracle.downsamp_depth(tot)
    ox, okept = oracle.downsample(m.indptr, m.data, n, 42)
    x = np.where(np.repeat(okept, np.diff(m.indptr)), ox, m.data).astype(np.uint32)
    Z, valid = oracle.features(m.indptr, m.indices, x, feat_map_of(m.ngenes, feats), feats.size)
    assert nv == valid.sum()
    ocol, _ = oracle.cor_knn(Z[valid], K * kx)
    deg, odst, _ = oracle.balance(ocol, K, kx, kb)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    cells = np.nonzero(valid)[0]
    got = set(zip(src.tolist(), dst.tolist()))
    exp = set(zip(cells[osrc].tolist(), cells[odst2].tolist()))
    jacc = len(got & exp) / len(got | exp)
    print(f"edges gpu {len(got)} oracle {len(exp)} jaccard {jacc:.6f}")
    assert jacc > 0.999          # differences only from correlation ties within 1e-5


def test_tgs_cor_graph_dense_entry(ctx, oracle):
    """tgs_cor_graph(x = feat, ...) on a dense matrix (reference R/utils.r:115) == CSC entry"""
    from metacell_b200 import api, _lib
    m = _small(ncells=800)
    feats = np.arange(0, 400, 2, dtype=np.int32)
    dense = m.to_dense()[feats].astype(np.float64)
    x = np.log2(1.0 + 7.0 * dense)                                     # R/cgraph_knn.r:13-14
    gr = api.tgs_cor_graph(x, 20, 10, 10, 3, ctx=ctx)
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(20, 10)
    g.mutual(3)
    g.prune()
    src, dst, w = g.edges()
    a = set(zip(gr["col1"].tolist(), gr["col2"].tolist()))
    b = set(zip(src.tolist(), dst.tolist()))
    assert len(a & b) / len(a | b) > 0.999
    g.destroy()


def test_errors_are_reported_not_thrown(ctx):
    from metacell_b200 import _lib
    m = _small(ncells=50)
    csc = _upload(ctx, m)
    with pytest.raises(_lib.McbError, match="MC-ERR"):
        _lib.CGraph.from_csc(ctx, csc, np.array([0, 0], np.int32), 7.0)          # duplicate feature row
    with pytest.raises(_lib.McbError, match="MC-ERR"):
        csc.downsample(0.5, 1)
    g = _lib.CGraph.from_csc(ctx, csc, np.arange(100, dtype=np.int32), 7.0)
    with pytest.raises(_lib.McbError):
        g.mutual(3)                                                                # before knn
    g.destroy()


@pytest.mark.parametrize("ncells,K,kx,kb", [(1200, 20, 10, 3), (5000, 100, 10, 3), (9000, 150, 10, 3), (300, 40, 10, 3)])
def test_routed_join_equals_direct_join(ctx, oracle, ncells, K, kx, kb):
    """the routed join (multi-rank default), the direct probing join and the blocked-list join (single-rank default) give
    bit-identical balanced graphs; several row blocks, Kc above and below the number of cells, self edges kept and dropped"""
    from metacell_b200 import _lib
    m = _small(ncells=ncells)
    feats = np.arange(0, 500, 2, dtype=np.int32)
    csc = _upload(ctx, m)
    out = {}
    for keep_self in (0, 1):
        for impl in (0, 1, 2):
            ctx.set_option("join_impl", impl)
            ctx.set_option("keep_self", keep_self)
            try:
                g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
                g.knn(K, kx)
                col, _ = g.download_knn()
                g.mutual(kb)
                g.prune()
                out[impl] = g.edges()
                g.destroy()
            finally:
                ctx.set_option("join_impl", -1)
                ctx.set_option("keep_self", 0)
        assert all(np.array_equal(a, b) for a, b in zip(out[0], out[1])), keep_self
        assert all(np.array_equal(a, b) for a, b in zip(out[2], out[1])), keep_self
        deg, odst, _ = oracle.balance(col, K, kx, kb, keep_self=bool(keep_self))
        osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
        assert np.array_equal(out[0][0], osrc) and np.array_equal(out[0][1], odst2) and np.array_equal(out[0][2], ow)
