"""GPU-vs-oracle parity AT THE BASELINE CONFIGS (BASELINE.json `configs`, SURVEY.md section 8: C1, C2, C3, C5), at
their own sizes, through the C ABI.

  C1  PBMC8k-shaped: 8,000 cells x 4,000 gene rows, 600 feature genes, K = 100, dsamp = T
      (vignettes/a-basic_pbmc8k.Rmd:128-146): the whole chain against the oracle chain
  C2  100,000 cells x 1,000 feature genes, K = 100: a row sample against the fp64 oracle + full-size invariants
  C3  the C1 graph, 16 of the 500 resamples (same host-supplied seeds and index sets), p = 0.75, min_mc_size = 20
  C5  Amphimedon-shaped (vignettes/d-amphimedon.Rmd:126-146): 20,000 low-UMI cells, K = 150, dsamp = F, with
      degenerate (constant feature vector) cells

Tolerances (north_star): correlations within 1e-5 absolute; kNN lists equal except inside runs tied within that
tolerance; balanced edges bit-exact on identical lists; down-sampled counts and co-occurrence counts bit-exact.
"""
import numpy as np
import pytest

from conftest import feat_map_of, knn_parity

pytestmark = pytest.mark.gpu

COR_TOL = 1e-5          # north_star: "correlations within 1e-5 absolute"


def _upload(ctx, m, data=None):
    from metacell_b200 import _lib
    return _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, (m.data if data is None else data).astype(np.float64))


def _edge_set(src, dst):
    return set(zip(np.asarray(src).tolist(), np.asarray(dst).tolist()))


# ---------------------------------------------------------------------------------------------------------------
# C1
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c1(ctx, oracle):
    from metacell_b200 import _lib, synth
    from oracle import oracle_np
    m, feats, prm = synth.config("C1")
    K, kx, kb, seed = prm["K"], 10, 3, 42
    assert m.ncells == 8000 and feats.size == 600 and K == 100 and prm["dsamp"]
    # --- GPU, staged so that every intermediate is visible
    csc = _upload(ctx, m)
    tot = csc.col_sums()
    depth = _lib.downsamp_depth(tot)
    ds = csc.downsample(depth, seed, add_non_dsamp=True)
    x_gpu, kept_gpu = ds.download()
    g = _lib.CGraph.from_csc(ctx, ds, feats, 7.0)
    z_gpu = g.features(feats.size)
    cells = g.valid_cells()
    g.knn(K, kx)
    col, cor = g.download_knn()
    stats = g.knn_stats()
    g.mutual(kb)
    g.prune()
    src, dst, w = g.edges()
    g.destroy()
    # --- oracle chain on the same input (R/cgraph_knn.r:10-25 through R/gset.r:151-214, R/utils.r:30-69,115-123)
    otot = m.col_sums()
    odepth = oracle.downsamp_depth(otot)
    ox, okept = oracle.downsample(m.indptr, m.data, odepth, seed)
    x_or = np.where(np.repeat(okept, np.diff(m.indptr)), ox, m.data).astype(np.uint32)     # add-back, R/gset.r:179-187
    Z, valid = oracle.features(m.indptr, m.indices, x_or, feat_map_of(m.ngenes, feats), feats.size)
    ocol, ocor = oracle_np.cor_knn_blas(Z[valid], K * kx + 8)
    return dict(m=m, feats=feats, K=K, kx=kx, kb=kb, seed=seed, tot=tot, otot=otot, depth=depth, odepth=odepth,
                x_gpu=x_gpu, kept_gpu=kept_gpu, x_or=x_or, okept=okept, z_gpu=z_gpu, Z=Z, valid=valid, cells=cells,
                col=col, cor=cor, ocol=ocol, ocor=ocor, stats=stats, src=src, dst=dst, w=w)


def test_c1_depth_and_downsample_bit_exact(c1):
    assert np.array_equal(c1["tot"], c1["otot"])
    assert c1["depth"] == c1["odepth"]
    assert np.array_equal(c1["kept_gpu"], c1["okept"])
    assert 0.5 < c1["okept"].mean() < 1.0                  # both branches (down-sampled and added back) are exercised
    assert np.array_equal(c1["x_gpu"].astype(np.uint32), c1["x_or"])
    m = c1["m"]
    cs = np.add.reduceat(np.concatenate([c1["x_gpu"], [0]]), m.indptr[:-1])[: m.ncells]
    assert np.all(cs[c1["okept"]] == np.floor(c1["depth"]))                     # R/utils.r:36-39


def test_c1_features_and_correlations(c1):
    assert np.array_equal(c1["cells"], np.nonzero(c1["valid"])[0])
    assert np.abs(c1["z_gpu"] - c1["Z"][c1["valid"]]).max() < 2e-7
    kc = c1["K"] * c1["kx"]
    err = np.abs(c1["cor"] - c1["ocor"][:, :kc]).max()
    print(f"C1: max |cor_gpu - cor_fp64| over {c1['cor'].size} kNN entries = {err:.2e}")
    assert err < COR_TOL


def test_c1_knn_lists_equal_off_ties(c1):
    agree = knn_parity(c1["col"], c1["cor"], c1["ocol"], c1["ocor"], COR_TOL)
    print(f"C1: kNN position agreement {agree:.6f}, filter stats {c1['stats']}")
    # knn_parity has already asserted that every disagreement sits inside a run tied within the tolerance; neighbouring
    # correlations of an 8,000-cell row lie ~3e-4 apart, so ~0.6 % of the positions are such near-ties
    assert agree > 0.99


def test_c1_balanced_edges_bit_exact_on_same_lists(c1, oracle):
    K, kx, kb = c1["K"], c1["kx"], c1["kb"]
    deg, odst, _ = oracle.balance(c1["col"], K, kx, kb)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    cells = c1["cells"]
    assert np.array_equal(c1["src"], cells[osrc]) and np.array_equal(c1["dst"], cells[odst2]) and np.array_equal(c1["w"], ow)
    n = c1["m"].ncells
    assert np.bincount(c1["src"], minlength=n).max() <= K and np.bincount(c1["dst"], minlength=n).max() <= K * kb
    assert abs(c1["src"].size / n - K) < 0.2 * K                                   # R/coclust_graph_cov.r:39


def test_c1_one_shot_entry_matches_oracle_chain(ctx, oracle, c1):
    """mcb_cgraph_bknn (the call the R glue makes) == oracle chain balanced from the ORACLE's own lists"""
    from metacell_b200 import _lib
    m, feats, K, kx, kb = c1["m"], c1["feats"], c1["K"], c1["kx"], c1["kb"]
    src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, K, kx, kb,
                                               7.0, dsamp=True, seed=c1["seed"])
    assert nv == c1["valid"].sum()
    assert np.array_equal(src, c1["src"]) and np.array_equal(dst, c1["dst"]) and np.array_equal(w, c1["w"])   # staged == one shot
    deg, odst, _ = oracle.balance(c1["ocol"][:, :K * kx], K, kx, kb)
    osrc, odst2, _ = oracle.edges_from_balance(deg, odst, K)
    cells = np.nonzero(c1["valid"])[0]
    got, exp = _edge_set(src, dst), _edge_set(cells[osrc], cells[odst2])
    jacc = len(got & exp) / len(got | exp)
    print(f"C1: edges gpu {len(got)} oracle {len(exp)} jaccard {jacc:.6f}")
    assert jacc > 0.999                                   # differences only from correlation ties within 1e-5


# ---------------------------------------------------------------------------------------------------------------
# C3: 16 of the 500 resamples of the C1 graph
# ---------------------------------------------------------------------------------------------------------------
def test_c3_sixteen_of_500_resamples_bit_exact(ctx, oracle, c1):
    from metacell_b200 import _lib, synth
    N = c1["m"].ncells
    src, dst, w = c1["src"], c1["dst"], c1["w"]
    sets, seeds = synth.resample_sets(N, 0.75, 500, 77)                # the 500 host-supplied sets of the bench's C3 leg
    pick = np.arange(0, 500, 32)[:16]
    assert pick.size == 16 and sets.shape == (500, 6000)
    cc = _lib.CoClust(ctx, N, src, dst, w, sets[pick], seeds[pick], 20, 1.05, 10)
    mc, sweeps = cc.assignments()
    n1, n2, cnt, ns = cc.pairs()
    cc.destroy()
    csr = oracle.graph_to_csr(N, src, dst, w)
    ocnt, ons, osw = oracle.coclust(N, csr, sets[pick], seeds[pick], 20, 1.05, 10)
    assert np.array_equal(sweeps, osw)
    for q in (0, 7, 15):
        omc, _ = oracle.graph_cover(N, csr, sets[pick[q]], 20, 1.05, 10, int(seeds[pick[q]]))
        assert np.array_equal(mc[q], omc), q
    dense = np.zeros((N, N), np.uint32)
    dense[n1, n2] = cnt
    assert np.array_equal(dense, ocnt) and np.array_equal(ns, ons)
    assert ns.sum() == 16 * 6000 and np.all(cnt <= np.minimum(ns[n1], ns[n2]))
    print(f"C3: 16 resamples, {n1.size} co-cluster pairs, sweeps {sweeps.min()}..{sweeps.max()}")


# ---------------------------------------------------------------------------------------------------------------
# C5-shaped: 20,000 low-UMI cells, K = 150, degenerate cells
# ---------------------------------------------------------------------------------------------------------------
def test_c5_shaped_low_umi_with_degenerate_cells(ctx, oracle):
    from metacell_b200 import _lib, synth
    from oracle import oracle_np
    m, feats, prm = synth.config("C5")
    K, kx, kb = prm["K"], 10, 3
    assert m.ncells == 20000 and K == 150 and not prm["dsamp"]
    # degenerate cells: no UMI in any feature gene (all-zero feature vector -> NA correlations, R/cgraph_knn_mcnorm.r:25)
    data = m.data.copy()
    fm = feat_map_of(m.ngenes, feats)
    rng = np.random.default_rng(5)
    dead = np.sort(rng.choice(m.ncells, 60, replace=False))
    for c in dead:
        sl = slice(m.indptr[c], m.indptr[c + 1])
        data[sl] = np.where(fm[m.indices[sl]] >= 0, 0, data[sl])
    csc = _upload(ctx, m, data)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    Z, valid = oracle.features(m.indptr, m.indices, data, fm, feats.size)
    assert not valid[dead].any() and valid.sum() <= m.ncells - 60
    cells = g.valid_cells()
    assert np.array_equal(cells, np.nonzero(valid)[0])
    g.knn(K, kx)
    col, cor = g.download_knn()
    stats = g.knn_stats()
    ocol, ocor = oracle_np.cor_knn_blas(Z[valid], K * kx + 8)
    agree = knn_parity(col, cor, ocol, ocor, COR_TOL)
    print(f"C5-shaped: {valid.sum()} valid cells, kNN agreement {agree:.6f}, stats {stats}")
    assert agree > 0.98
    g.mutual(kb)
    g.prune()
    src, dst, w = g.edges()
    g.destroy()
    deg, odst, _ = oracle.balance(col, K, kx, kb)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    assert np.array_equal(src, cells[osrc]) and np.array_equal(dst, cells[odst2]) and np.array_equal(w, ow)
    touched = np.zeros(m.ncells, bool)
    touched[src] = True
    touched[dst] = True
    assert not touched[~valid].any()                       # "sim graph is missing N nodes" path, R/cgraph.r:36-39


# ---------------------------------------------------------------------------------------------------------------
# C2: the bench workload at its own size -- a row sample against fp64 + size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_c2_row_sample_vs_fp64_and_invariants(ctx, oracle):
    from metacell_b200 import _lib, synth
    from oracle import oracle_np
    m, feats, prm = synth.config("C2")
    K, kx, kb = prm["K"], 10, 3
    assert m.ncells == 100000 and feats.size == 1000 and K == 100
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    cells = g.valid_cells()
    g.knn(K, kx)
    col, cor = g.download_knn()
    stats = g.knn_stats()
    g.mutual(kb)
    g.prune()
    src, dst, w = g.edges()
    g.destroy()
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    assert np.array_equal(cells, np.nonzero(valid)[0])
    Zv = Z[valid]
    M = Zv.shape[0]
    rows = np.sort(np.random.default_rng(2).choice(M, 256, replace=False))
    ocol, ocor = oracle_np.cor_knn_blas(Zv, K * kx + 8, rows=rows)
    agree = knn_parity(col[rows], cor[rows], ocol, ocor, COR_TOL)
    err = np.abs(cor[rows] - ocor[:, :K * kx]).max()
    print(f"C2: 256 rows of {M}: max |dcor| {err:.2e}, kNN agreement {agree:.6f}, stats {stats}")
    assert agree > 0.97 and err < COR_TOL
    # size-independent properties of the full result
    assert np.array_equal(col[:, 0], np.arange(M)) and np.all(cor[:, 0] == 1.0)        # self at rank 1, R/atlas.r:475-476
    assert np.all(np.diff(cor[:, 1:], axis=1) <= 0)                                     # ranked by descending correlation
    srt = np.sort(col, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1]) and srt.min() >= 0 and srt.max() < M       # distinct, in range
    assert np.bincount(src, minlength=m.ncells).max() <= K and np.bincount(dst, minlength=m.ncells).max() <= K * kb
    assert np.all(src != dst)
    first = np.concatenate([[True], src[1:] != src[:-1]])
    assert np.all(w[first] == 1.0) and w.min() > 0.0                                     # best edge of every source has w = 1
    # balancing at full size, bit-exact on the GPU's own lists
    deg, odst, _ = oracle.balance(col, K, kx, kb)
    osrc, odst2, ow = oracle.edges_from_balance(deg, odst, K)
    assert np.array_equal(src, cells[osrc]) and np.array_equal(dst, cells[odst2]) and np.array_equal(w, ow)


# ---------------------------------------------------------------------------------------------------------------
# the fused sampling pass (corsample.cu) on a hard shape: low UMI counts, wide size factors, rare types -- rows whose
# correlation distribution is far from the bulk's must still get thresholds that land in their candidate window
# ---------------------------------------------------------------------------------------------------------------
def test_fused_sampling_on_rare_types_and_low_umi(ctx, oracle):
    from metacell_b200 import _lib, synth
    from oracle import oracle_np
    m = synth.make_umi_matrix(45000, 500, n_types=40, median_umis=600.0, size_sigma=0.6, rare_types=8, seed=77)
    feats = np.arange(500, dtype=np.int32)
    K, kx = 100, 10
    csc = _upload(ctx, m)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g.knn(K, kx)
    col, cor = g.download_knn()
    stats = g.knn_stats()
    cells = g.valid_cells()
    g.destroy()
    assert stats["sample_cols"] >= 4096                                    # the fused path was taken
    assert stats["fallback_rows"] <= 45, stats                              # <= 0.1 % of the rows needed the exact path
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    assert np.array_equal(cells, np.nonzero(valid)[0])
    Zv = Z[valid]
    rare = np.flatnonzero(m.cell_type[valid] < 8)                           # every cell of the rare types + a random sample
    rows = np.unique(np.concatenate([rare[:200], np.random.default_rng(3).choice(Zv.shape[0], 200, replace=False)]))
    ocol, ocor = oracle_np.cor_knn_blas(Zv, K * kx + 8, rows=rows)
    agree = knn_parity(col[rows], cor[rows], ocol, ocor, COR_TOL)
    print(f"rare types / low UMI: {rows.size} rows ({min(200, rare.size)} of rare types), kNN agreement {agree:.6f}, stats {stats}")
    assert agree > 0.97
