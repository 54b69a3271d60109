"""GPU parity of the resampled co-clustering stage (A8) against the CPU oracle: per-resample cluster
assignments, sweep counts and the integer co-occurrence matrix must be bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _graph(oracle, ncells=900, K=15, seed=5):
    from metacell_b200 import synth
    from conftest import feat_map_of
    m = synth.make_umi_matrix(ncells, 400, n_types=6, median_umis=2000.0, seed=seed)
    feats = np.arange(0, 400, 2, dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    assert valid.all()
    col, _ = oracle.cor_knn(Z, K * 10)
    deg, dst, _ = oracle.balance(col, K)
    return oracle.edges_from_balance(deg, dst, K)


@pytest.mark.parametrize("ncells,K,min_size,n_resamp", [(900, 15, 8, 6), (1500, 30, 20, 4)])
def test_cover_assignments_and_counts_bit_exact(ctx, oracle, ncells, K, min_size, n_resamp):
    from metacell_b200 import _lib, synth
    src, dst, w = _graph(oracle, ncells, K)
    sets, seeds = synth.resample_sets(ncells, 0.75, n_resamp, 11)
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10)
    mc, sweeps = cc.assignments()
    csr = oracle.graph_to_csr(ncells, src, dst, w)
    for r in range(n_resamp):
        omc, osw = oracle.graph_cover(ncells, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert sweeps[r] == osw, (r, sweeps[r], osw)
        assert np.array_equal(mc[r], omc), r
    n1, n2, cnt, ns = cc.pairs()
    ocnt, ons, _ = oracle.coclust(ncells, csr, sets, seeds, min_size, 1.05, 10)
    dense = np.zeros((ncells, ncells), np.uint32)
    dense[n1, n2] = cnt
    assert np.array_equal(dense, ocnt)
    assert np.array_equal(ns, ons)
    # invariants
    assert np.all(n1 < n2) and np.all(cnt > 0)
    assert ns.sum() == n_resamp * sets.shape[1] and ns.max() <= n_resamp
    assert np.all(cnt <= np.minimum(ns[n1], ns[n2]))
    order = np.lexsort((n2, n1))
    assert np.array_equal(order, np.arange(n1.size))               # sorted output
    cc.destroy()


@pytest.mark.parametrize("kernel", ["auto", "incr8", "incr16", "batch2", "batch4", "slow"])
@pytest.mark.parametrize("ncells,K,knn,min_size", [(900, 15, 9, 8), (1400, 30, 20, 15)])
def test_knn_edge_truncation_bit_exact(ctx, oracle, monkeypatch, ncells, K, knn, min_size, kernel):
    """`knn` of tgs_graph_cover_resample (R/coclust_graph_cov.r:39-41): every resample's cover must equal the oracle's cover
    of the resample's truncated graph (oracle.truncate_edges: at most knn highest-weight sampled out edges per node)."""
    from metacell_b200 import _lib, synth
    monkeypatch.setenv("MCB_COVER_KERNEL", kernel)
    src, dst, w = _graph(oracle, ncells, K)
    n_resamp = 4
    sets, seeds = synth.resample_sets(ncells, 0.75, n_resamp, 23)
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10, knn=knn)
    mc, sweeps = cc.assignments()
    free = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10, knn=0)
    mc_free, _ = free.assignments()
    free.destroy()
    dense_o = np.zeros((ncells, ncells), np.uint32)
    n_trunc = 0
    for r in range(n_resamp):
        ts, td, tw = oracle.truncate_edges(ncells, src, dst, w, sets[r], knn)
        inside = np.zeros(ncells, bool); inside[sets[r]] = True
        n_trunc += int((inside[src] & inside[dst]).sum()) - ts.size
        omc, osw = oracle.graph_cover(ncells, oracle.graph_to_csr(ncells, ts, td, tw), sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert sweeps[r] == osw, (r, sweeps[r], osw)
        assert np.array_equal(mc[r], omc), r
        for k in range(omc.max() + 1):
            mem = np.flatnonzero(omc == k)
            dense_o[np.ix_(mem, mem)] += 1
    assert n_trunc > 0, "the case must actually truncate"
    assert not np.array_equal(mc, mc_free), "truncation changed nothing: the case does not exercise it"
    n1, n2, cnt, ns = cc.pairs()
    dense = np.zeros((ncells, ncells), np.uint32)
    dense[n1, n2] = cnt
    assert np.array_equal(dense, np.triu(dense_o, 1))
    cc.destroy()


def test_knn_at_or_above_max_out_degree_is_no_limit(ctx, oracle):
    from metacell_b200 import _lib, synth
    ncells, K = 700, 12
    src, dst, w = _graph(oracle, ncells, K)
    sets, seeds = synth.resample_sets(ncells, 0.75, 3, 4)
    a = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, knn=0)
    b = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, knn=K)
    assert all(np.array_equal(x, y) for x, y in zip(a.assignments(), b.assignments()))
    a.destroy(); b.destroy()


@pytest.mark.parametrize("knn", [0, 9])
def test_method_edges_counts_equal_the_full_matrix_on_graph_edges(ctx, oracle, knn):
    """method = "edges": one counter per graph edge; equals the N x N matrix of method = "full" read at the edges, and the
    oracle's covers"""
    from metacell_b200 import _lib, synth
    ncells, K, min_size, n_resamp = 900, 15, 8, 6
    src, dst, w = _graph(oracle, ncells, K)
    sets, seeds = synth.resample_sets(ncells, 0.75, n_resamp, 11)
    ce = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10, knn=knn, method="edges")
    e1, e2, ec, ens = ce.edge_counts()
    cf = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10, knn=knn)
    n1, n2, cnt, ns = cf.pairs()
    dense = np.zeros((ncells, ncells), np.int64)
    dense[n1, n2] = cnt
    dense = dense + dense.T
    assert e1.size == src.size and set(zip(e1.tolist(), e2.tolist())) == set(zip(src.tolist(), dst.tolist()))
    assert np.array_equal(ec, dense[e1, e2]) and np.array_equal(ens, ns)
    wmap = {(a, b): x for a, b, x in zip(src.tolist(), dst.tolist(), w.tolist())}
    wfix = np.maximum(1, np.rint(np.array([wmap[(a, b)] for a, b in zip(e1.tolist(), e2.tolist())]) * 65536.0)).astype(np.int64)
    assert np.array_equal(np.lexsort((e2, -wfix, e1)), np.arange(e1.size))      # rows ordered by (source, weight descending, target)
    # against the oracle's own covers (no knn: the whole graph)
    if knn == 0:
        csr = oracle.graph_to_csr(ncells, src, dst, w)
        want = np.zeros(e1.size, np.int64)
        for r in range(n_resamp):
            omc, _ = oracle.graph_cover(ncells, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
            want += (omc[e1] >= 0) & (omc[e1] == omc[e2])
        assert np.array_equal(ec, want)
    ce.destroy(); cf.destroy()


def test_method_edges_needs_no_n_squared_memory(ctx):
    """300,000 nodes: method = "full" is refused (N x N counters), method = "edges" runs"""
    from metacell_b200 import _lib
    N = 300000
    rng = np.random.default_rng(2)
    base = np.arange(N, dtype=np.int32)
    src = np.concatenate([base, base, base])
    dst = np.concatenate([(base + 1) % N, (base + 2) % N, (base // 64 * 64 + rng.integers(0, 64, N)).astype(np.int32) % N]).astype(np.int32)
    keep = src != dst
    key = src[keep].astype(np.int64) * N + dst[keep]
    _, first = np.unique(key, return_index=True)
    src, dst = src[keep][first], dst[keep][first]
    w = rng.random(src.size) * 0.9 + 0.1
    sets = np.sort(np.stack([rng.choice(N, int(0.75 * N), replace=False) for _ in range(2)]), axis=1).astype(np.int32)
    seeds = np.array([5, 6], np.uint64)
    with pytest.raises(_lib.McbError, match="N x N"):
        _lib.CoClust(ctx, N, src, dst, w, sets, seeds, 3, 1.05, 10)
    cc = _lib.CoClust(ctx, N, src, dst, w, sets, seeds, 3, 1.05, 10, method="edges")
    e1, e2, ec, ns = cc.edge_counts()
    mc, _ = cc.assignments()
    want = sum(((mc[r][e1] >= 0) & (mc[r][e1] == mc[r][e2])).astype(np.int64) for r in range(2))
    assert np.array_equal(ec, want) and ec.max() <= 2 and ec.sum() > 0
    assert ns.sum() == 2 * sets.shape[1]
    cc.destroy()


def test_rank_sharding_sums_to_whole(ctx, oracle):
    """resamples strided over 3 emulated ranks: counters add up to the single-rank result"""
    from metacell_b200 import _lib, synth
    ncells = 700
    src, dst, w = _graph(oracle, ncells, 12)
    sets, seeds = synth.resample_sets(ncells, 0.75, 7, 2)
    whole = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10)
    n1, n2, cnt, ns = whole.pairs()
    dense = np.zeros((ncells, ncells), np.int64)
    dense[n1, n2] = cnt
    acc = np.zeros_like(dense)
    accs = np.zeros(ncells, np.int64)
    for r in range(3):
        part = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, rank=r, nranks=3)
        a, b, c, s = part.pairs()
        acc[a, b] += c
        accs += s
        part.destroy()
    assert np.array_equal(acc, dense) and np.array_equal(accs, ns)
    whole.destroy()


def test_api_mirror_coclust(ctx, oracle, tmp_path):
    """mcell_coclust_from_graph_resamp through the host mirror (reference R/coclust_graph_cov.r:13-58)"""
    from metacell_b200 import api
    ncells = 600
    src, dst, w = _graph(oracle, ncells, 10)
    api.scdb_init(str(tmp_path), force_reinit=True)
    api.scdb_add_cgraph("g", api.tgCellGraph({"mc1": src, "mc2": dst, "w": w}, [f"c{i}" for i in range(ncells)]))
    coc = api.mcell_coclust_from_graph_resamp("coc", "g", 8, 0.75, 5, ctx=ctx)
    assert api.scdb_coclust("coc") is coc
    assert coc.n_samp.sum() == 5 * round(0.75 * ncells)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mcell_coclust_from_graph_resamp("coc2", "missing", 8, 0.75, 5, ctx=ctx)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgCoClust("missing", coc.coclust, coc.n_samp)


def test_single_graph_cover_and_mc_from_coclust(ctx, oracle, tmp_path):
    """tgs_graph_cover (reference R/mc_graphcov.r:27) == one oracle cover over all nodes, and the two callers after the
    path (mcell_add_mc_from_graph, mcell_mc_from_coclust_balanced) run through the mirror"""
    from metacell_b200 import api, _lib
    ncells = 700
    src, dst, w = _graph(oracle, ncells, 12)
    cluster, sweeps = _lib.graph_cover(ctx, ncells, src, dst, w, 8, 1.05, 10, 42)
    csr = oracle.graph_to_csr(ncells, src, dst, w)
    omc, osw = oracle.graph_cover(ncells, csr, np.arange(ncells, dtype=np.int32), 8, 1.05, 10, 42)
    assert sweeps == osw and np.array_equal(cluster, np.where(omc < 0, 0, omc + 1))
    api.scdb_init(str(tmp_path), force_reinit=True)
    cells = [f"c{i}" for i in range(ncells)]
    api.scdb_add_mat("m", api.tgScMat([0] * (ncells + 1), [], [], genes=["g"], cells=cells))
    api.scdb_add_cgraph("g", api.tgCellGraph({"mc1": src, "mc2": dst, "w": w}, cells))
    res = api.mcell_add_mc_from_graph("mc1", "g", "m", 8, ctx=ctx)
    assert len(res["mc"]) + len(res["outliers"]) == ncells
    assert sorted(set(res["mc"].values())) == list(range(1, max(res["mc"].values()) + 1))      # as.integer(as.factor())
    coc = api.mcell_coclust_from_graph_resamp("coc", "g", 8, 0.75, 20, ctx=ctx)
    filt = api.mcell_coclust_filt_by_k_deg("coc", 10, 2)
    assert filt.dtype == bool and filt.size == coc.coclust["cnt"].size and 0 < filt.sum() <= filt.size
    res2 = api.mcell_mc_from_coclust_balanced("mc2", "coc", "m", 10, 8, alpha=2, ctx=ctx)
    assert len(res2["mc"]) + len(res2["outliers"]) == ncells and len(res2["mc"]) > ncells // 2
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mcell_mc_from_coclust_balanced("mc3", "nope", "m", 10, 8, ctx=ctx)


@pytest.mark.parametrize("kernel", ["incr8", "incr16", "batch2", "batch4", "batch8", "warp8", "warp16", "warp32", "fast", "slow"])
def test_graph_cover_kernel_variants_bit_exact(ctx, oracle, kernel, monkeypatch):
    """Every cover kernel (incremental per-node cluster tables, speculative 4-warp batches, speculative warp-per-node,
    sequential shared-memory, sequential global-memory) evaluates the same Gauss-Seidel pass: assignments and sweep counts must equal the
    oracle's for each of them (MCB_COVER_KERNEL forces the choice inside launch_graph_cover)."""
    from metacell_b200 import _lib, synth
    ncells, min_size = 1100, 10
    src, dst, w = _graph(oracle, ncells, 20, seed=9)
    monkeypatch.setenv("MCB_COVER_KERNEL", kernel)
    sets, seeds = synth.resample_sets(ncells, 0.75, 5, 4242)
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10)
    mc, sweeps = cc.assignments()
    cc.destroy()
    csr = oracle.graph_to_csr(ncells, src, dst, w)
    for r in range(sets.shape[0]):
        omc, osw = oracle.graph_cover(ncells, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert sweeps[r] == osw, (kernel, r, sweeps[r], osw)
        assert np.array_equal(mc[r], omc), (kernel, r)


@pytest.mark.parametrize("kernel", ["auto", "incr8", "incr16", "batch2", "batch4", "warp16", "warp8", "fast", "slow"])
def test_graph_cover_high_degree_graph(ctx, oracle, kernel, monkeypatch):
    """K = 140 graph (out-degree > 128, in-degree up to 3K as at K = 150 in C4/C5): the wider-slot instantiations
    of the speculative kernels and the 256-thread sequential kernel against the oracle"""
    from metacell_b200 import _lib, synth
    ncells, min_size = 1500, 25
    src, dst, w = _graph(oracle, ncells, 140, seed=13)
    assert np.bincount(src).max() > 128
    monkeypatch.setenv("MCB_COVER_KERNEL", kernel)
    sets, seeds = synth.resample_sets(ncells, 0.75, 3, 99)
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10)
    mc, sweeps = cc.assignments()
    cc.destroy()
    csr = oracle.graph_to_csr(ncells, src, dst, w)
    for r in range(sets.shape[0]):
        omc, osw = oracle.graph_cover(ncells, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert sweeps[r] == osw, (kernel, r, sweeps[r], osw)
        assert np.array_equal(mc[r], omc), (kernel, r)


def test_duplicate_edges_are_refused(ctx, oracle):
    """(source, target) pairs are unique in every cell graph and the cover kernels rely on it (one thread per edge claims a
    seed's neighbours): an edge table that repeats a pair is an MC-ERR, not a silently different cover"""
    from metacell_b200 import _lib, synth
    ncells = 800
    src, dst, w = _graph(oracle, ncells, 14, seed=21)
    extra = np.random.default_rng(3).choice(src.size, 5, replace=False)
    src2, dst2, w2 = np.concatenate([src, src[extra]]), np.concatenate([dst, dst[extra]]), np.concatenate([w, w[extra] * 0.5])
    sets, seeds = synth.resample_sets(ncells, 0.75, 2, 77)
    with pytest.raises(_lib.McbError, match="MC-ERR.*more than once"):
        _lib.CoClust(ctx, ncells, src2, dst2, w2, sets, seeds, 8, 1.05, 10)
    _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10).destroy()


def test_cover_of_a_graph_beyond_shared_memory(ctx, oracle):
    """100,000 nodes, thousands of clusters: the per-resample state no longer fits shared memory; the resampled cover and
    tgs_graph_cover (general kernel: assignment in global memory) equal the oracle's"""
    from metacell_b200 import _lib
    N, min_size = 100000, 8
    rng = np.random.default_rng(8)
    base = np.arange(N, dtype=np.int64)
    blk = base // 64 * 64
    src = np.concatenate([base] * 12)
    dst = np.concatenate([(base + 1) % N, (base + 2) % N] + [blk + rng.integers(0, 64, N) for _ in range(10)]) % N
    keep = src != dst
    _, first = np.unique(src[keep] * N + dst[keep], return_index=True)
    src, dst = src[keep][first].astype(np.int32), dst[keep][first].astype(np.int32)
    w = rng.random(src.size) * 0.9 + 0.1
    sets = np.sort(np.stack([rng.choice(N, int(0.75 * N), replace=False) for _ in range(2)]), axis=1).astype(np.int32)
    seeds = np.array([15, 16], np.uint64)
    cc = _lib.CoClust(ctx, N, src, dst, w, sets, seeds, min_size, 1.05, 10, method="edges")
    mc, sweeps = cc.assignments()
    cc.destroy()
    csr = oracle.graph_to_csr(N, src, dst, w)
    for r in range(2):
        omc, osw = oracle.graph_cover(N, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert sweeps[r] == osw and np.array_equal(mc[r], omc), r
        assert omc.max() > 1000                                    # thousands of clusters
    cluster, nsw = _lib.graph_cover(ctx, N, src, dst, w, min_size, 1.05, 10, 99)
    omc, osw = oracle.graph_cover(N, csr, np.arange(N, dtype=np.int32), min_size, 1.05, 10, 99)
    assert nsw == osw and np.array_equal(cluster, np.where(omc < 0, 0, omc + 1))


@pytest.mark.parametrize("kernel", ["auto", "incr8", "incr16"])
def test_incremental_cover_with_more_clusters_than_one_load_round(ctx, oracle, kernel, monkeypatch):
    """more than 128 clusters: the incremental kernel reads a node's cluster rows in several rounds of 128 (the C1 / C3-sized
    cases stay inside one); assignments, sweep counts and counts against the oracle"""
    from metacell_b200 import _lib, synth
    ncells, K, min_size = 3600, 12, 6
    src, dst, w = _graph(oracle, ncells, K, seed=31)
    monkeypatch.setenv("MCB_COVER_KERNEL", kernel)
    sets, seeds = synth.resample_sets(ncells, 0.75, 3, 5)
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, min_size, 1.05, 10)
    mc, sweeps = cc.assignments()
    cc.destroy()
    csr = oracle.graph_to_csr(ncells, src, dst, w)
    for r in range(sets.shape[0]):
        omc, osw = oracle.graph_cover(ncells, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
        assert omc.max() + 1 > 128, "the case must need a second round of row loads"
        assert sweeps[r] == osw, (kernel, r, sweeps[r], osw)
        assert np.array_equal(mc[r], omc), (kernel, r)
