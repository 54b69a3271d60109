"""Host <-> device staging of the objects that cross the C ABI (pageable R vectors in, edge table out), the input
validation that rides on it, and the entry points added for the advisor's findings: the single graph cover without
N x N counters, the `knn` argument of the resampled cover, contexts on the same device used from several threads."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _small(ncells=900, ngenes=400, seed=3, **kw):
    from metacell_b200 import synth
    return synth.make_umi_matrix(ncells, ngenes, n_types=6, median_umis=2000.0, seed=seed, **kw)


def _upload(ctx, m, data=None, indices=None, indptr=None, ngenes=None):
    from metacell_b200 import _lib
    return _lib.Csc.upload(ctx, m.ngenes if ngenes is None else ngenes, m.indptr if indptr is None else indptr,
                           m.indices if indices is None else indices, (m.data if data is None else data).astype(np.float64))


def test_upload_round_trip_and_escaped_counts(ctx):
    """counts >= 65535 leave the 16-bit staging format through the patch list; the device copy equals the host data"""
    m = _small()
    data = m.data.astype(np.float64)
    data[5] = 65535.0
    data[77] = 70000.0
    data[m.nnz - 1] = 4294967295.0
    csc = _upload(ctx, m, data)
    x, kept = csc.download()
    assert np.array_equal(x, data) and kept.all()
    tot = csc.col_sums()
    exp = np.add.reduceat(np.concatenate([data, [0]]), m.indptr[:-1])[: m.ncells]
    assert np.array_equal(tot, exp.astype(np.int64))


def test_upload_wide_row_indices(ctx):
    """more than 65536 gene rows: the row indices stay 32-bit on the way up"""
    m = _small(ncells=300)
    idx = m.indices.astype(np.int64) * 200 + 7            # still ascending within a column
    csc = _upload(ctx, m, indices=idx.astype(np.int32), ngenes=400 * 200 + 8)
    from metacell_b200 import _lib
    feats = (np.arange(0, 400, 2) * 200 + 7).astype(np.int32)
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    g2 = _lib.CGraph.from_csc(ctx, _upload(ctx, m), np.arange(0, 400, 2, dtype=np.int32), 7.0)
    assert np.array_equal(g.features(feats.size), g2.features(feats.size))
    g.destroy(); g2.destroy()


@pytest.mark.parametrize("case", ["fraction", "negative", "too_big", "nan", "p_order", "row_range", "row_order", "row_dup"])
def test_malformed_matrices_are_rejected(ctx, case):
    from metacell_b200 import _lib
    m = _small(ncells=200)
    data, idx, ptr = m.data.astype(np.float64), m.indices.copy(), m.indptr.copy()
    ngenes = m.ngenes
    if case == "fraction": data[10] = 1.5
    if case == "negative": data[10] = -1.0
    if case == "too_big": data[10] = 4294967296.0
    if case == "nan": data[10] = np.nan
    if case == "p_order": ptr[50], ptr[51] = ptr[51], ptr[50]
    if case == "row_range": ngenes = int(idx.max())                    # ngenes smaller than max(i) + 1
    if case == "row_order":
        a = ptr[3]
        idx[a], idx[a + 1] = idx[a + 1], idx[a]
    if case == "row_dup": idx[ptr[3] + 1] = idx[ptr[3]]
    with pytest.raises(_lib.McbError, match="MC-ERR"):
        _upload(ctx, m, data, idx, ptr, ngenes)


def test_upload_range_is_the_column_slice(ctx):
    from metacell_b200 import _lib
    m = _small()
    whole = _upload(ctx, m).col_sums()
    for a, b in ((0, 300), (300, 301), (301, 900), (450, 450)):
        part = _lib.Csc.upload_range(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), a, b)
        assert np.array_equal(part.col_sums(), whole[a:b])
        part.free()


def test_staged_edge_download_equals_device_table(ctx):
    """(mc1, mc2, w) rebuilt by the host threads from targets + out-degrees == the table the device emitted"""
    import ctypes as C
    from metacell_b200 import _lib
    m = _small(ncells=2500)
    g = _lib.CGraph.from_csc(ctx, _upload(ctx, m), np.arange(0, 400, 2, dtype=np.int32), 7.0)
    g.knn(30, 10); g.mutual(3); ne = g.prune()
    src, dst, w = g.edges()
    n = C.c_int64(0)
    ps, pd, pw = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(_lib.lib().mcb_cgraph_edges_device(g.h, C.byref(n), C.byref(ps), C.byref(pd), C.byref(pw)))
    assert n.value == ne
    import torch

    class _DevMem:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    ctx.sync()
    dev = [torch.as_tensor(_DevMem(p.value, ne * sz), device=torch.device("cuda", ctx.device)).view(t).cpu().numpy()
           for t, p, sz in zip((torch.int32, torch.int32, torch.float64), (ps, pd, pw), (4, 4, 8))]
    assert np.array_equal(src, dev[0]) and np.array_equal(dst, dev[1]) and np.array_equal(w, dev[2])
    g.destroy()


def test_host_thread_option_gives_identical_results(ctx):
    from metacell_b200 import _lib
    m = _small(ncells=1500)
    feats = np.arange(0, 400, 2, dtype=np.int32)
    ref = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, 20, 10, 3, 7.0)
    for nt in (1, 3):
        ctx.set_option("host_threads", nt)
        got = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, 20, 10, 3, 7.0)
        assert all(np.array_equal(a, b) for a, b in zip(got[:3], ref[:3]))
    ctx.set_option("host_threads", 0)


def test_two_contexts_on_one_device_from_two_threads(oracle):
    """the library keeps no process-wide mutable state on the launch path (per-device opt-ins, thread-local errors)"""
    from metacell_b200 import _lib
    m = _small(ncells=1200)
    feats = np.arange(0, 400, 2, dtype=np.int32)
    x = m.data.astype(np.float64)
    out, err = [None, None], []

    def work(q):
        try:
            c = _lib.Context(0)
            out[q] = _lib.cgraph_bknn_oneshot(c, m.ngenes, m.indptr, m.indices, x, feats, 20, 10, 3, 7.0)
            c.close()
        except Exception as e:          # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(q,)) for q in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    assert all(np.array_equal(a, b) for a, b in zip(out[0][:3], out[1][:3]))


def test_graph_cover_of_a_large_sparse_graph_needs_no_counter_matrix(ctx, oracle):
    """tgs_graph_cover on 300,000 nodes (N*N counters would be 360 GB): disjoint cliques of 25 nodes"""
    from metacell_b200 import _lib
    N, c = 300000, 25
    base = (np.arange(N) // c) * c
    src = np.repeat(np.arange(N, dtype=np.int32), c - 1)
    offs = np.tile(np.arange(1, c), N)
    dst = (base.repeat(c - 1) + (np.arange(N).repeat(c - 1) % c + offs) % c).astype(np.int32)
    w = np.tile(1.0 - np.arange(c - 1) / 32.0, N)
    cluster, sweeps = _lib.graph_cover(ctx, N, src, dst, w, 20, 1.05, 10, 99)
    csr = oracle.graph_to_csr(N, src, dst, w)
    omc, osw = oracle.graph_cover(N, csr, np.arange(N, dtype=np.int32), 20, 1.05, 10, 99)
    assert sweeps == osw and np.array_equal(cluster, np.where(omc < 0, 0, omc + 1))
    assert (cluster > 0).mean() > 0.99
    # every clique lands in one cluster
    assert np.all(cluster.reshape(-1, c).min(axis=1) == cluster.reshape(-1, c).max(axis=1))


def test_coclust_knn_argument(ctx, oracle):
    from metacell_b200 import _lib, synth
    from test_gpu_coclust import _graph
    ncells = 900
    src, dst, w = _graph(oracle, ncells, 15)
    sets, seeds = synth.resample_sets(ncells, 0.75, 5, 11)
    ref = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10).pairs()
    max_out = int(np.bincount(src).max())
    for knn in (max_out, max_out + 7):                      # no node can exceed it: identical to "no limit"
        got = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, knn=knn).pairs()
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    # below the largest out-degree the limit binds: edges are truncated (parity with the oracle's truncated graphs is in
    # test_gpu_coclust.py::test_knn_edge_truncation_bit_exact), never silently ignored
    cut = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, knn=3).pairs()
    assert not all(np.array_equal(a, b) for a, b in zip(cut, ref))
    with pytest.raises(_lib.McbError, match="knn"):
        _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 8, 1.05, 10, knn=-1)
