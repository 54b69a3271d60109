"""CPU tests of the oracle (oracle/mc_oracle.c): external known answers where they exist (Philox KAT,
base-R quantile/round semantics), independent numpy / pure-Python restatements of each stage on small
inputs, the invariants the reference's R code implies, and the committed golden
vectors.  PARITY UNPINNED: the reference has no golden vectors of its own for this path."""
import os

import numpy as np
import pytest

from conftest import feat_map_of

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz")


def _mat(ncells=240, ngenes=120, seed=99, **kw):
    from metacell_b200 import synth
    return synth.make_umi_matrix(ncells, ngenes, n_types=4, median_umis=900.0, seed=seed, **kw)


# ---- RNG: published Random123 known-answer vectors for Philox4x32-10 ---------------------------------
@pytest.mark.parametrize("ctr,key,exp", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat(oracle, ctr, key, exp):
    assert [int(v) for v in oracle.philox(ctr, key)] == exp


# ---- A1 ----------------------------------------------------------------------------------------------
def test_depth_rule_matches_r_semantics(oracle):
    rng = np.random.default_rng(1)
    for n in (5, 17, 400, 4001):
        tot = rng.integers(200, 9000, size=n)
        q50 = np.round(np.quantile(tot, 0.5))            # numpy default = R type 7; np.round = half-even
        q05 = np.round(np.quantile(tot, 0.05))
        assert oracle.downsamp_depth(tot) == min(q50, max(750.0, q05))
    assert oracle.downsamp_depth(np.array([1, 2])) == 2.0      # median 1.5 rounds half-even to 2
    assert oracle.downsamp_depth(np.array([2, 3])) == 2.0      # 2.5 -> 2


# ---- A2 ----------------------------------------------------------------------------------------------
def test_downsample_invariants(oracle):
    m = _mat()
    tot = m.col_sums()
    n = 600.5
    ds, kept = oracle.downsample(m.indptr, m.data, n, 42)
    assert np.array_equal(kept, tot >= n)                                   # R/utils.r:34 (raw n)
    cs = np.array([ds[m.indptr[c]:m.indptr[c + 1]].sum() for c in range(m.ncells)])
    assert np.all(cs[kept] == 600) and np.all(cs[~kept] == 0)              # depth = floor(n)
    assert np.all(ds <= m.data)
    ds2, _ = oracle.downsample(m.indptr, m.data, n, 42)
    assert np.array_equal(ds, ds2)                                          # deterministic
    ds3, _ = oracle.downsample(m.indptr, m.data, n, 43)
    assert not np.array_equal(ds, ds3)                                      # seed matters


def test_downsample_is_uniform_without_replacement(oracle):
    """one cell, many seeds: per-gene mean follows the hypergeometric expectation d * v_g / tot"""
    v = np.array([50, 30, 15, 4, 1], dtype=np.uint32)
    indptr = np.array([0, 5], dtype=np.int64)
    d = 40
    acc = np.zeros(5)
    reps = 4000
    for s in range(reps):
        out, kept = oracle.downsample(indptr, v, float(d), s)
        assert kept[0] and out.sum() == d
        acc += out
    exp = d * v / v.sum()
    sd = np.sqrt(d * (v / 100) * (1 - v / 100) * (100 - d) / 99 / reps)
    assert np.all(np.abs(acc / reps - exp) < 5 * sd + 1e-9)


def test_downsample_python_restatement(oracle):
    """pure-Python statement of the key rule on a tiny cell"""
    v = np.array([3, 0, 2, 4], dtype=np.uint32)
    indptr = np.array([0, 4], dtype=np.int64)
    seed, c, d = 12345, 0, 5
    keys = []
    for t in range(int(v.sum())):
        o = oracle.philox([t >> 2, c, 0, 0x4453], [seed & 0xffffffff, seed >> 32])
        keys.append((int(o[t & 3]) << 32) | t)
    thr = sorted(keys)[d - 1]
    exp, t = [], 0
    for x in v:
        exp.append(sum(1 for q in range(x) if keys[t + q] <= thr))
        t += int(x)
    out, _ = oracle.downsample(indptr, v, float(d), seed)
    assert out.tolist() == exp


# ---- A3 / A4 -----------------------------------------------------------------------------------------
def test_features_vs_numpy(oracle):
    m = _mat()
    feats = np.array([7, 3, 50, 51, 52, 90, 11], dtype=np.int32)            # gene-set order, not sorted
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    X = np.log2(1.0 + 7.0 * m.to_dense()[feats].astype(np.float64)).T       # R/cgraph_knn.r:13-14
    Xc = X - X.mean(axis=1, keepdims=True)
    ss = (Xc ** 2).sum(axis=1)
    ok = ss >= 1e-20
    assert np.array_equal(valid, ok)
    assert np.allclose(Z[ok], Xc[ok] / np.sqrt(ss[ok])[:, None], atol=1e-14)
    assert np.all(Z[~ok] == 0)
    # Pearson: Z Z^T == np.corrcoef
    assert np.allclose((Z[ok] @ Z[ok].T), np.corrcoef(X[ok]), atol=1e-12)


# ---- A5 ----------------------------------------------------------------------------------------------
def test_cor_knn_vs_bruteforce(oracle):
    from oracle import oracle_np
    m = _mat()
    feats = np.arange(0, 120, 2, dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    Zv = Z[valid]
    col, cor = oracle.cor_knn(Zv, 40)
    C = Zv @ Zv.T
    for i in range(Zv.shape[0]):
        c = C[i].copy()
        c[i] = 2.0
        order = np.lexsort((np.arange(c.size), -c))[:40]
        assert np.array_equal(col[i], order)
    assert np.all(col[:, 0] == np.arange(Zv.shape[0])) and np.all(cor[:, 0] == 1.0)      # R/atlas.r:475-476
    assert np.all(np.diff(cor[:, 1:], axis=1) <= 1e-15)
    c2, r2 = oracle_np.cor_knn_blas(Zv, 40)                       # dgemm + orc_select_rows (OpenMP)
    assert np.array_equal(col, c2) and np.abs(cor - r2).max() < 1e-12
    c3, r3 = oracle_np.cor_knn_numpy(Zv, 40)                      # numpy-only selection
    assert np.array_equal(col, c3) and np.abs(cor - r3).max() < 1e-12
    pick = np.array([7, 3, 50, 51, 52, 11], dtype=np.int64)       # explicit row lists (the bench's parity sample)
    c4, r4 = oracle_np.cor_knn_blas(Zv, 40, rows=pick)
    assert np.array_equal(c4, col[pick]) and np.abs(r4 - cor[pick]).max() < 1e-12
    # Kc > M pads with -1; row ranges agree with the full call
    colp, _ = oracle.cor_knn(Zv[:30], 40)
    assert np.all(colp[:, 30:] == -1) and np.all(colp[:, :30] >= 0)
    cs, _ = oracle.cor_knn(Zv, 40, 10, 25)
    assert np.array_equal(cs, col[10:25])


# ---- A6 ----------------------------------------------------------------------------------------------
def _balance_py(col, K, k_expand, k_beta, strict=True, keep_self=False):
    """dict-based restatement of SURVEY App. A6 (rule of R/cgraph_knn_mcnorm.r:48-81)"""
    M, Kc = col.shape
    rank = [{int(j): r + 1 for r, j in enumerate(col[i]) if j >= 0} for i in range(M)]
    smax = K * K * k_expand
    pairs = {}
    for i in range(M):
        for j, rij in rank[i].items():
            if i == j and not keep_self:
                continue
            rji = rank[j].get(i)
            if rji is None:
                continue
            s = rij * rji
            if (s < smax) if strict else (s <= smax):
                pairs[(i, j)] = s
    keep = set()
    for j in range(M):
        inc = sorted((s, i) for (i, jj), s in pairs.items() if jj == j)[: K * k_beta]
        keep |= {(i, j) for s, i in inc}
    out = []
    for i in range(M):
        o = sorted((pairs[(i, j)], j) for (ii, j) in keep if ii == i)[:K]
        out.append(o)
    return out


@pytest.mark.parametrize("K,kx,kb", [(4, 3, 2), (6, 5, 3), (3, 10, 1)])
def test_balance_vs_python_restatement(oracle, K, kx, kb):
    m = _mat(ncells=90)
    feats = np.arange(0, 120, 2, dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    col, _ = oracle.cor_knn(Z[valid], K * kx)
    for strict in (True, False):
        for keep_self in (False, True):
            deg, dst, s = oracle.balance(col, K, kx, kb, thresh_strict=strict, keep_self=keep_self)
            exp = _balance_py(col, K, kx, kb, strict, keep_self)
            for i in range(col.shape[0]):
                assert deg[i] == len(exp[i])
                assert [(int(s[i, q]), int(dst[i, q])) for q in range(deg[i])] == exp[i]


def test_balance_invariants(oracle):
    m = _mat()
    feats = np.arange(0, 120, 2, dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    K, kx, kb = 8, 5, 3
    col, _ = oracle.cor_knn(Z[valid], K * kx)
    deg, dst, s = oracle.balance(col, K, kx, kb)
    src, d, w = oracle.edges_from_balance(deg, dst, K)
    M = col.shape[0]
    assert deg.max() <= K and np.bincount(d, minlength=M).max() <= K * kb
    assert np.all(s[dst >= 0] < K * K * kx) and np.all(s[dst >= 0] >= 1)
    assert np.all(src != d)                                                  # self edges dropped
    assert w.max() == 1.0 and w.min() > 0 and np.allclose(w * K, np.rint(w * K))
    first = np.r_[True, src[1:] != src[:-1]]
    assert np.all(w[first] == 1.0)                                           # best edge of each source has w = 1
    # mutuality: every kept edge's reverse direction was in the Kc lists
    lists = [set(col[i][col[i] >= 0].tolist()) for i in range(M)]
    assert all(int(a) in lists[int(b)] and int(b) in lists[int(a)] for a, b in zip(src, d))


# ---- A7 / A8 -----------------------------------------------------------------------------------------
def _graph(oracle, ncells=240, K=8):
    m = _mat(ncells=ncells)
    feats = np.arange(0, 120, 2, dtype=np.int32)
    Z, valid = oracle.features(m.indptr, m.indices, m.data, feat_map_of(m.ngenes, feats), feats.size)
    col, _ = oracle.cor_knn(Z[valid], K * 5)
    deg, dst, _ = oracle.balance(col, K, 5, 3)
    return int(valid.sum()), oracle.edges_from_balance(deg, dst, K)


def test_graph_cover_properties(oracle):
    from metacell_b200 import synth
    N, (src, dst, w) = _graph(oracle)
    csr = oracle.graph_to_csr(N, src, dst, w)
    sets, seeds = synth.resample_sets(N, 0.75, 4, 5)
    for r in range(4):
        mc, sweeps = oracle.graph_cover(N, csr, sets[r], 5, 1.05, 10, int(seeds[r]))
        sampled = np.zeros(N, bool)
        sampled[sets[r]] = True
        assert np.all(mc[~sampled] == -2) and np.all(mc[sampled] >= -1)      # cluster 0 / outlier semantics
        assert 1 <= sweeps <= 1000
        ids = np.unique(mc[mc >= 0])
        assert ids.size >= 2 and ids.max() < N
        mc2, sw2 = oracle.graph_cover(N, csr, sets[r], 5, 1.05, 10, int(seeds[r]))
        assert np.array_equal(mc, mc2) and sweeps == sw2
    mc_a, _ = oracle.graph_cover(N, csr, sets[0], 5, 1.05, 10, 1)
    mc_b, _ = oracle.graph_cover(N, csr, sets[0], 5, 1.05, 10, 2)
    assert not np.array_equal(mc_a, mc_b)
    # a node with no sampled neighbour stays an outlier; empty graph -> everything outlier
    empty = oracle.graph_to_csr(N, np.zeros(0, int), np.zeros(0, int), np.zeros(0))
    mc_e, _ = oracle.graph_cover(N, empty, sets[0], 5, 1.05, 10, 1)
    assert np.all(mc_e[sets[0]] == -1)


def test_coclust_invariants(oracle):
    from metacell_b200 import synth
    N, (src, dst, w) = _graph(oracle)
    csr = oracle.graph_to_csr(N, src, dst, w)
    n_resamp = 5
    sets, seeds = synth.resample_sets(N, 0.75, n_resamp, 5)
    cnt, nsamp, sweeps = oracle.coclust(N, csr, sets, seeds, 5, 1.05, 10)
    assert nsamp.sum() == n_resamp * sets.shape[1] and nsamp.max() <= n_resamp
    assert np.all(np.tril(cnt) == 0)                                                     # a < b only, no diagonal
    a, b = np.nonzero(cnt)
    assert np.all(cnt[a, b] <= np.minimum(nsamp[a], nsamp[b]))
    # equals the sum of per-resample same-cluster indicator matrices
    acc = np.zeros((N, N), np.uint32)
    for r in range(n_resamp):
        mc, _ = oracle.graph_cover(N, csr, sets[r], 5, 1.05, 10, int(seeds[r]))
        same = (mc[:, None] == mc[None, :]) & (mc[:, None] >= 0)
        acc += np.triu(same, 1).astype(np.uint32)
    assert np.array_equal(acc, cnt)


def test_incremental_cover_structure_equals_the_oracle(oracle, tmp_path):
    """tests/cover_incr_model.c is the host model of what graph_cover_incr_kernel keeps (per-node cluster rows edited on
    moves instead of re-read edges): same assignments and sweep counts as orc_graph_cover, most visits are not moves
    (the premise of the kernel), and an fp32 evaluation of the scores decides the arg-max correctly whenever the runner-up
    is more than 1e-5 behind (the kernel's shortcut; closer calls are scored in fp64)"""
    import ctypes as C
    import subprocess
    from metacell_b200 import synth
    so = str(tmp_path / "libcover_incr_model.so")
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(["gcc", "-O2", "-mavx2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(here, "cover_incr_model.c"), "-lm"])
    L = C.CDLL(so)

    def p(a, t):
        return a.ctypes.data_as(C.POINTER(t))

    for ncells, K, min_size in ((240, 8, 5), (700, 20, 10)):
        N, (src, dst, w) = _graph(oracle, ncells, K)
        optr, odst, ow, iptr, isrc, iw = csr = oracle.graph_to_csr(N, src, dst, w)
        sets, seeds = synth.resample_sets(N, 0.75, 3, 5)
        for r in range(3):
            mc, sweeps = oracle.graph_cover(N, csr, sets[r], min_size, 1.05, 10, int(seeds[r]))
            mc2 = np.zeros(N, np.int32); nsw = C.c_int32(0); moves = np.zeros(1000, np.int64); f32 = np.zeros(3, np.int64)
            smp = np.ascontiguousarray(sets[r], dtype=np.int32)
            rc = L.model_graph_cover_incr(C.c_int32(N), p(optr, C.c_int64), p(odst, C.c_int32), p(ow, C.c_uint32), p(iptr, C.c_int64),
                                          p(isrc, C.c_int32), p(iw, C.c_uint32), C.c_int32(smp.size), p(smp, C.c_int32),
                                          C.c_int32(min_size), C.c_double(1.05), C.c_int32(10), C.c_uint64(int(seeds[r])),
                                          C.c_int32(1000), p(mc2, C.c_int32), C.byref(nsw), p(moves, C.c_int64), p(f32, C.c_int64))
            assert rc == 0
            # the kernel's fp32 pre-decision: never wrong when it clears its 1e-5 margin, and it clears it almost always
            assert f32[0] > 0 and f32[2] == 0 and f32[1] < 0.02 * f32[0], f32
            assert nsw.value == sweeps and np.array_equal(mc, mc2)
            assert moves[:sweeps].sum() < 0.5 * sweeps * smp.size and moves[sweeps - 1] == 0


# ---- golden vectors -------------------------------------------------------------------------------------
def test_golden_vectors(oracle):
    g = np.load(GOLDEN)
    ngenes = int(g["ngenes"])
    feats = g["feats"]
    tot = np.add.reduceat(g["data"].astype(np.int64), g["indptr"][:-1])
    assert oracle.downsamp_depth(tot) == float(g["depth"])
    ds, kept = oracle.downsample(g["indptr"], g["data"], 600.0, 42)
    assert np.array_equal(ds, g["ds"]) and np.array_equal(kept, g["kept"])
    Z, valid = oracle.features(g["indptr"], g["indices"], g["data"], feat_map_of(ngenes, feats), feats.size)
    assert np.array_equal(valid, g["valid"]) and np.abs(Z - g["Z"]).max() < 1e-14
    col, cor = oracle.cor_knn(Z[valid], 40)
    assert np.array_equal(col, g["knn_col"]) and np.abs(cor - g["knn_cor"]).max() < 1e-13
    deg, dst, s = oracle.balance(col, 8, 5, 3)
    assert np.array_equal(deg, g["deg"]) and np.array_equal(dst, g["bal_dst"]) and np.array_equal(s, g["bal_s"])
    src_e, dst_e, w_e = oracle.edges_from_balance(deg, dst, 8)
    M = int(valid.sum())
    csr = oracle.graph_to_csr(M, src_e, dst_e, w_e)
    cnt, nsamp, sweeps = oracle.coclust(M, csr, g["sets"], g["seeds"], 5, 1.05, 10)
    assert np.array_equal(sweeps, g["sweeps"]) and np.array_equal(nsamp, g["nsamp"])
    n1, n2 = np.nonzero(cnt)
    assert np.array_equal(n1, g["pair1"]) and np.array_equal(n2, g["pair2"]) and np.array_equal(cnt[n1, n2], g["pair_cnt"])
    for r in range(3):
        mc, _ = oracle.graph_cover(M, csr, g["sets"][r], 5, 1.05, 10, int(g["seeds"][r]))
        assert np.array_equal(mc, g["mcs"][r])


def test_truncate_edges_keeps_the_knn_heaviest_sampled_out_edges():
    """oracle.truncate_edges (`knn` of tgs_graph_cover_resample, R/coclust_graph_cov.r:39-41) against a literal loop"""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    N, knn = 60, 4
    src, dst = np.nonzero(rng.random((N, N)) < 0.2)
    keep = src != dst
    src, dst = src[keep].astype(np.int32), dst[keep].astype(np.int32)
    w = np.round(rng.random(src.size) * 8) / 8 + 1.0 / 65536                   # many tied weights
    sample = np.sort(rng.choice(N, 45, replace=False)).astype(np.int32)
    ts, td, tw = O.truncate_edges(N, src, dst, w, sample, knn)
    inside = set(sample.tolist())
    want = set()
    for v in inside:
        cand = [(-(max(1, int(np.rint(w[e] * 65536.0)))), int(dst[e])) for e in np.flatnonzero(src == v) if int(dst[e]) in inside]
        for _, u in sorted(cand)[:knn]:
            want.add((v, u))
    assert set(zip(ts.tolist(), td.tolist())) == want
    a, b, c = O.truncate_edges(N, src, dst, w, sample, 0)
    assert a.size == sum(1 for s_, d_ in zip(src.tolist(), dst.tolist()) if s_ in inside and d_ in inside)


def test_device_form_of_the_knn_rule_equals_truncate_edges():
    """The cover kernels never build a resample's truncated edge table: out segments are stored by (weight descending,
    target ascending), a resample keeps ocut[v] = the leading positions of v's segment up to (not including) its
    (knn+1)-th sampled target, and the edge u -> v is usable iff its position in u's segment is below ocut[u]
    (coclust.cu: the `ocut` loop of every kernel's set-up, `ipos[e] < ocut[source]`).  Restated here literally and compared
    with oracle.truncate_edges, the form the oracle's covers are run on."""
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    for N, dens, knn in ((80, 0.25, 5), (120, 0.15, 3), (50, 0.5, 40)):
        src, dst = np.nonzero(rng.random((N, N)) < dens)
        keep = src != dst
        src, dst = src[keep].astype(np.int32), dst[keep].astype(np.int32)
        w = np.round(rng.random(src.size) * 6) / 6 + 1.0 / 65536               # many tied weights
        sample = np.sort(rng.choice(N, int(0.75 * N), replace=False)).astype(np.int32)
        sampled = np.zeros(N, bool); sampled[sample] = True
        wf = np.maximum(1, np.rint(w * 65536.0)).astype(np.int64)
        usable = set()
        for v in range(N):
            e = np.flatnonzero(src == v)
            e = e[np.lexsort((dst[e], -wf[e]))]                                # the stored order of v's out segment
            c, ocut = 0, e.size
            for pos, ee in enumerate(e):                                      # the kernels' ocut loop
                if sampled[dst[ee]]:
                    if c == knn:
                        ocut = pos
                        break
                    c += 1
            if not sampled[v]:
                continue                                                      # unsampled nodes have no ocut: none of their edges is ever read
            for pos, ee in enumerate(e):
                if pos < ocut and sampled[dst[ee]]:                           # targets outside the sample are skipped at use
                    usable.add((int(v), int(dst[ee])))
        ts, td, _ = O.truncate_edges(N, src, dst, w, sample, knn)
        assert usable == set(zip(ts.tolist(), td.tolist())), (N, knn)
