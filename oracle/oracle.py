"""ctypes view of the CPU oracle (oracle/mc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under metacell_b200/ imports this module.
PARITY UNPINNED (see mc_oracle.c header): the reference holds no golden vectors for this path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # built with -mavx2 (no -march=native) so the object travels to the GPU box
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = C.CDLL(build())
            _lib.orc_num_threads()
        except Exception:  # e.g. built with -march=native on another host
            _lib = C.CDLL(build(force=True))
        _lib.orc_downsamp_depth.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads():
    return int(lib().orc_num_threads())


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def downsamp_depth(tot):
    tot = np.ascontiguousarray(tot, dtype=np.int64)
    return float(lib().orc_downsamp_depth(_p(tot, C.c_int64), C.c_int64(tot.size)))


def downsample(indptr, data, n, seed):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    data = np.ascontiguousarray(data, dtype=np.uint32)
    ncells = indptr.size - 1
    out = np.zeros_like(data)
    kept = np.zeros(ncells, dtype=np.uint8)
    rc = lib().orc_downsample(C.c_int64(ncells), _p(indptr, C.c_int64), _p(data, C.c_uint32), C.c_double(n),
                              C.c_uint64(seed), _p(out, C.c_uint32), _p(kept, C.c_uint8))
    if rc:
        raise RuntimeError(f"orc_downsample rc={rc}")
    return out, kept.astype(bool)


def features(indptr, indices, data, feat_map, G, k_nonz=7.0):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.uint32)
    feat_map = np.ascontiguousarray(feat_map, dtype=np.int32)
    ncells = indptr.size - 1
    Z = np.zeros((ncells, G), dtype=np.float64)
    valid = np.zeros(ncells, dtype=np.uint8)
    rc = lib().orc_features(C.c_int64(ncells), _p(indptr, C.c_int64), _p(indices, C.c_int32), _p(data, C.c_uint32),
                            _p(feat_map, C.c_int32), C.c_int32(G), C.c_double(k_nonz), _p(Z, C.c_double),
                            _p(valid, C.c_uint8))
    if rc:
        raise RuntimeError(f"orc_features rc={rc}")
    return Z, valid.astype(bool)


def cor_knn(Z, Kc, row_begin=0, row_end=None):
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    M, G = Z.shape
    row_end = M if row_end is None else row_end
    nr = row_end - row_begin
    col = np.full((nr, Kc), -1, dtype=np.int32)
    cor = np.zeros((nr, Kc), dtype=np.float64)
    rc = lib().orc_cor_knn(C.c_int64(M), C.c_int32(G), _p(Z, C.c_double), C.c_int32(Kc), C.c_int64(row_begin),
                           C.c_int64(row_end), _p(col, C.c_int32), _p(cor, C.c_double))
    if rc:
        raise RuntimeError(f"orc_cor_knn rc={rc}")
    return col, cor


def select_rows(Cblock, self_cols, Kc):
    """orc_select_rows: per row of a dense fp64 correlation block the best min(Kc, M) columns, (cor desc, col asc),
    self forced to rank 1; all host cores (OpenMP)."""
    Cblock = np.ascontiguousarray(Cblock, dtype=np.float64)
    nr, M = Cblock.shape
    sc = np.ascontiguousarray(self_cols, dtype=np.int64)
    col = np.empty((nr, Kc), dtype=np.int32)
    cor = np.empty((nr, Kc), dtype=np.float64)
    rc = lib().orc_select_rows(C.c_int64(nr), C.c_int64(M), _p(Cblock, C.c_double), C.c_int64(M), _p(sc, C.c_int64),
                               C.c_int32(Kc), _p(col, C.c_int32), _p(cor, C.c_double))
    if rc:
        raise RuntimeError(f"orc_select_rows rc={rc}")
    return col, cor


def balance(knn_col, K, k_expand=10, k_beta=3, thresh_strict=True, keep_self=False):
    knn_col = np.ascontiguousarray(knn_col, dtype=np.int32)
    M, Kc = knn_col.shape
    deg = np.zeros(M, dtype=np.int32)
    dst = np.full((M, K), -1, dtype=np.int32)
    s = np.zeros((M, K), dtype=np.int32)
    rc = lib().orc_balance(C.c_int64(M), C.c_int32(Kc), _p(knn_col, C.c_int32), C.c_int32(K), C.c_int32(k_expand),
                           C.c_int32(k_beta), C.c_int(int(thresh_strict)), C.c_int(int(keep_self)),
                           _p(deg, C.c_int32), _p(dst, C.c_int32), _p(s, C.c_int32))
    if rc:
        raise RuntimeError(f"orc_balance rc={rc}")
    return deg, dst, s


def edges_from_balance(deg, dst, K):
    """(src, dst, w) with w = 1 - (place-1)/K, ordered by (src, place)."""
    M = deg.size
    place = np.arange(dst.shape[1])[None, :]
    mask = place < deg[:, None]
    src = np.broadcast_to(np.arange(M, dtype=np.int32)[:, None], dst.shape)[mask]
    w = np.broadcast_to(1.0 - place / float(K), dst.shape)[mask]
    return src.astype(np.int32), dst[mask].astype(np.int32), w.astype(np.float64)


def graph_to_csr(N, src, dst, w):
    """Out- and in-CSR with 16-bit fixed-point weights (mc_oracle.c A7 DECISION `weights`)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    wf = np.maximum(1, np.rint(np.asarray(w, dtype=np.float64) * 65536.0)).astype(np.uint32)
    o = np.lexsort((dst, src))
    optr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(optr, src + 1, 1)
    optr = np.cumsum(optr)
    i = np.lexsort((src, dst))
    iptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(iptr, dst + 1, 1)
    iptr = np.cumsum(iptr)
    return (optr, dst[o].astype(np.int32), wf[o], iptr, src[i].astype(np.int32), wf[i])


def truncate_edges(N, src, dst, w, sample, knn):
    """`knn` of tgs_graph_cover_resample (reference call site R/coclust_graph_cov.r:39-41, which passes the graph's average
    out-degree): inside one resample a node uses at most `knn` out edges.  DECISION `knn_rule` (tgstat is absent): among a
    sampled node's out edges whose target is sampled too, keep the `knn` with the highest weight (16-bit fixed point as in
    graph_to_csr; ties go to the lower target index); the dropped edges do not exist for that resample, in either
    direction.  Returns the edge table of the resample's graph (edges with an unsampled endpoint are removed as well;
    orc_graph_cover ignores them anyway).  knn <= 0: no limit."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    inside = np.zeros(N, dtype=bool)
    inside[np.asarray(sample)] = True
    keep = inside[src] & inside[dst]
    src, dst, w = src[keep], dst[keep], w[keep]
    if knn is None or knn <= 0 or src.size == 0:
        return src.astype(np.int32), dst.astype(np.int32), w
    wf = np.maximum(1, np.rint(w * 65536.0)).astype(np.int64)
    order = np.lexsort((dst, -wf, src))                          # by source, weight descending, target ascending
    s_o = src[order]
    first = np.r_[0, np.flatnonzero(np.diff(s_o)) + 1]
    start = np.repeat(first, np.diff(np.r_[first, s_o.size]))
    place = np.arange(s_o.size) - start                           # position among the node's sampled out edges
    sel = order[place < knn]
    sel.sort()
    return src[sel].astype(np.int32), dst[sel].astype(np.int32), w[sel]


def graph_cover(N, csr, sample, min_size, cooling, burn_in, seed, max_sweeps=1000):
    optr, odst, ow, iptr, isrc, iw = csr
    sample = np.ascontiguousarray(sample, dtype=np.int32)
    mc = np.zeros(N, dtype=np.int32)
    nsw = C.c_int32(0)
    rc = lib().orc_graph_cover(C.c_int32(N), _p(optr, C.c_int64), _p(odst, C.c_int32), _p(ow, C.c_uint32),
                               _p(iptr, C.c_int64), _p(isrc, C.c_int32), _p(iw, C.c_uint32),
                               C.c_int32(sample.size), _p(sample, C.c_int32), C.c_int32(min_size),
                               C.c_double(cooling), C.c_int32(burn_in), C.c_uint64(seed), C.c_int32(max_sweeps),
                               _p(mc, C.c_int32), C.byref(nsw))
    if rc:
        raise RuntimeError(f"orc_graph_cover rc={rc}")
    return mc, int(nsw.value)


def coclust(N, csr, samples, seeds, min_size, cooling, burn_in, max_sweeps=1000):
    optr, odst, ow, iptr, isrc, iw = csr
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    n_resamp, n_s = samples.shape
    cnt = np.zeros((N, N), dtype=np.uint32)
    nsamp = np.zeros(N, dtype=np.int32)
    sweeps = np.zeros(n_resamp, dtype=np.int32)
    rc = lib().orc_coclust(C.c_int32(N), _p(optr, C.c_int64), _p(odst, C.c_int32), _p(ow, C.c_uint32),
                           _p(iptr, C.c_int64), _p(isrc, C.c_int32), _p(iw, C.c_uint32),
                           C.c_int32(n_resamp), C.c_int32(n_s), _p(samples, C.c_int32), _p(seeds, C.c_uint64),
                           C.c_int32(min_size), C.c_double(cooling), C.c_int32(burn_in), C.c_int32(max_sweeps),
                           _p(cnt, C.c_uint32), _p(nsamp, C.c_int32), _p(sweeps, C.c_int32))
    if rc:
        raise RuntimeError(f"orc_coclust rc={rc}")
    return cnt, nsamp, sweeps
