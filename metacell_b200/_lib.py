"""ctypes binding of libmcb200.so (include/mcb200.h).

The library is the product: if it is missing or no B200 is visible, every call raises -- there is
no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcb200.so")

# every symbol include/mcb200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "mcb_version", "mcb_last_error", "mcb_launch_count", "mcb_ctx_create", "mcb_ctx_destroy", "mcb_ctx_sync", "mcb_ctx_stream",
    "mcb_ctx_set_option", "mcb_ctx_timings", "mcb_ctx_timings_reset", "mcb_downsamp_depth", "mcb_csc_upload",
    "mcb_csc_wrap_device", "mcb_csc_free", "mcb_csc_col_sums", "mcb_csc_downsample", "mcb_csc_download",
    "mcb_cgraph_create", "mcb_cgraph_create_dense", "mcb_cgraph_destroy", "mcb_cgraph_num_valid",
    "mcb_cgraph_valid_cells", "mcb_cgraph_features", "mcb_cgraph_knn", 
    
    "mcb_cgraph_download_knn", "mcb_cgraph_knn_stats", "mcb_cgraph_mutual", 
    "mcb_cgraph_prune", "mcb_cgraph_edges", "mcb_cgraph_debug_cor", "mcb_cor_knn_xy", "mcb_gene_stats", "mcb_mc_gene_tapply", "mcb_cgraph_bknn", "mcb_free",
    "mcb_coclust_run", "mcb_graph_cover", "mcb_coclust_destroy", "mcb_coclust_buffers", "mcb_coclust_assignments",
    "mcb_coclust_extract", "mcb_coclust_pairs",
    "mcb_csc_upload_range", "mcb_cgraph_edges_device", "mcb_multi_create", "mcb_multi_destroy", "mcb_multi_size", "mcb_multi_ctx",
    "mcb_multi_set_option", "mcb_cgraph_bknn_multi", "mcb_coclust_multi", "mcb_comm_unique_id", "mcb_comm_create",
    "mcb_comm_destroy", "mcb_comm_rank", "mcb_cgraph_build_sharded", "mcb_cgraph_gather_edges", "mcb_coclust_reduce",
    "mcb_coclust_run2", "mcb_coclust_run_edges", "mcb_coclust_edge_counts",
]


class McbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise McbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.mcb_last_error.restype = C.c_char_p
        L.mcb_ctx_stream.restype = C.c_void_p
        L.mcb_launch_count.restype = C.c_int64
        L.mcb_multi_ctx.restype = C.c_void_p
        for name in ("mcb_ctx_destroy", "mcb_csc_free", "mcb_cgraph_destroy", "mcb_coclust_destroy", "mcb_free", "mcb_comm_destroy",
                     "mcb_multi_destroy"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def launch_count():
    return int(lib().mcb_launch_count())


def check(rc):
    if rc != 0:
        raise McbError(lib().mcb_last_error().decode("utf-8", "replace"))


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """mcb_ctx: one device + stream + memory pool."""

    def __init__(self, device=0, borrowed=None):
        self.owned = borrowed is None
        if borrowed is None:
            self.h = C.c_void_p()
            check(lib().mcb_ctx_create(C.c_int(device), C.byref(self.h)))
        else:
            self.h = C.c_void_p(borrowed)           # e.g. a device context of a Multi session
        self.device = device

    def close(self):
        if self.h and self.owned:
            lib().mcb_ctx_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().mcb_ctx_sync(self.h))

    def stream(self):
        return lib().mcb_ctx_stream(self.h)

    def set_option(self, key, value):
        check(lib().mcb_ctx_set_option(self.h, key.encode(), C.c_int64(int(value))))

    def timings(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        launches = (C.c_int64 * cap)()
        n = C.c_int(0)
        check(lib().mcb_ctx_timings(self.h, cap, names, ms, launches, C.byref(n)))
        return {names[i].decode(): (float(ms[i]), int(launches[i])) for i in range(min(n.value, cap))}

    def timings_reset(self):
        check(lib().mcb_ctx_timings_reset(self.h))


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def downsamp_depth(col_sums):
    cs = as_c(col_sums, np.int64)
    d = C.c_double(0.0)
    check(lib().mcb_downsamp_depth(ptr(cs, C.c_int64), C.c_int64(cs.size), C.byref(d)))
    return float(d.value)


class Csc:
    """mcb_csc: device-resident genes x cells count matrix."""

    def __init__(self, ctx, handle, ngenes, ncells, nnz, keep=None):
        self.ctx, self.h, self.ngenes, self.ncells, self.nnz = ctx, handle, ngenes, ncells, nnz
        self._keep = keep            # objects that must outlive this handle (shared structure / wrapped tensors)

    @classmethod
    def upload(cls, ctx, ngenes, indptr, indices, data):
        p = as_c(indptr, np.int64)
        i = as_c(indices, np.int32)
        x = as_c(data, np.float64)
        h = C.c_void_p()
        check(lib().mcb_csc_upload(ctx.h, C.c_int32(ngenes), C.c_int64(p.size - 1), ptr(p, C.c_int64),
                                   ptr(i, C.c_int32), ptr(x, C.c_double), C.byref(h)))
        return cls(ctx, h, ngenes, p.size - 1, int(p[-1]))

    @classmethod
    def upload_range(cls, ctx, ngenes, indptr, indices, data, cell0, cell1):
        """cells [cell0, cell1) of the host matrix: what one rank of a sharded build uploads (mcb_csc_upload_range)"""
        p = as_c(indptr, np.int64)
        i = as_c(indices, np.int32)
        x = as_c(data, np.float64)
        h = C.c_void_p()
        pp = C.cast(C.c_void_p(p.ctypes.data + 8 * int(cell0)), C.POINTER(C.c_int64))
        check(lib().mcb_csc_upload_range(ctx.h, C.c_int32(ngenes), C.c_int64(int(cell1) - int(cell0)), pp, ptr(i, C.c_int32),
                                         ptr(x, C.c_double), C.byref(h)))
        return cls(ctx, h, ngenes, int(cell1) - int(cell0), int(p[cell1] - p[cell0]))

    @classmethod
    def upload_ptrs(cls, ctx, ngenes, ncells, nnz, p_ptr, i_ptr, x_ptr, keep=None):
        """host pointers given as integers (e.g. pinned torch tensors' data_ptr())"""
        h = C.c_void_p()
        check(lib().mcb_csc_upload(ctx.h, C.c_int32(ngenes), C.c_int64(ncells), C.c_void_p(p_ptr), C.c_void_p(i_ptr),
                                   C.c_void_p(x_ptr), C.byref(h)))
        return cls(ctx, h, ngenes, ncells, nnz, keep)

    @classmethod
    def wrap_device(cls, ctx, ngenes, ncells, nnz, p_ptr, i_ptr, x_ptr, keep=None):
        h = C.c_void_p()
        check(lib().mcb_csc_wrap_device(ctx.h, C.c_int32(ngenes), C.c_int64(ncells), C.c_void_p(p_ptr),
                                        C.c_void_p(i_ptr), C.c_void_p(x_ptr), C.byref(h)))
        return cls(ctx, h, ngenes, ncells, nnz, keep)

    def free(self):
        if self.h:
            lib().mcb_csc_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def col_sums(self):
        tot = np.zeros(self.ncells, dtype=np.int64)
        check(lib().mcb_csc_col_sums(self.ctx.h, self.h, ptr(tot, C.c_int64)))
        return tot

    def downsample(self, n, seed, add_non_dsamp=False):
        h = C.c_void_p()
        check(lib().mcb_csc_downsample(self.ctx.h, self.h, C.c_double(n), C.c_uint64(int(seed)),
                                       C.c_int(int(add_non_dsamp)), C.byref(h)))
        return Csc(self.ctx, h, self.ngenes, self.ncells, self.nnz, keep=self)

    def download(self):
        x = np.zeros(self.nnz, dtype=np.float64)
        kept = np.zeros(self.ncells, dtype=np.uint8)
        check(lib().mcb_csc_download(self.ctx.h, self.h, ptr(x, C.c_double), ptr(kept, C.c_uint8)))
        return x, kept.astype(bool)


class CGraph:
    """mcb_cgraph: one balanced-graph build."""

    def __init__(self, ctx, handle, keep=None):
        self.ctx, self.h, self._keep = ctx, handle, keep
        m = C.c_int64(0)
        check(lib().mcb_cgraph_num_valid(self.h, C.byref(m)))
        self.M = int(m.value)
        self.rows = self.Kc = self.K = None

    @classmethod
    def from_csc(cls, ctx, csc, feat_rows, k_nonz=7.0):
        fr = as_c(feat_rows, np.int32)
        h = C.c_void_p()
        check(lib().mcb_cgraph_create(ctx.h, csc.h, ptr(fr, C.c_int32), C.c_int32(fr.size), C.c_double(k_nonz),
                                      C.byref(h)))
        return cls(ctx, h, keep=csc)

    @classmethod
    def from_dense(cls, ctx, x_cells_by_genes):
        """x: [ncells][G] float64 C-contiguous == R's genes x cells column-major matrix"""
        x = as_c(x_cells_by_genes, np.float64)
        h = C.c_void_p()
        check(lib().mcb_cgraph_create_dense(ctx.h, ptr(x, C.c_double), C.c_int32(x.shape[1]), C.c_int64(x.shape[0]),
                                            C.byref(h)))
        return cls(ctx, h)

    def destroy(self):
        if self.h:
            lib().mcb_cgraph_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def valid_cells(self):
        out = np.zeros(self.M, dtype=np.int32)
        check(lib().mcb_cgraph_valid_cells(self.h, ptr(out, C.c_int32)))
        return out

    def features(self, G):
        z = np.zeros((self.M, G), dtype=np.float32)
        check(lib().mcb_cgraph_features(self.h, ptr(z, C.c_float)))
        return z

    def debug_cor(self, impl, a_rows, b_rows):
        c = np.zeros((a_rows, b_rows), dtype=np.float32)
        check(lib().mcb_cgraph_debug_cor(self.h, C.c_int(impl), C.c_int64(a_rows), C.c_int64(b_rows), ptr(c, C.c_float)))
        return c

    def knn(self, K, k_expand=10, rank=0, nranks=1):
        check(lib().mcb_cgraph_knn(self.h, C.c_int32(K), C.c_int32(k_expand), C.c_int32(rank), C.c_int32(nranks)))
        self._set_split(K, k_expand, rank, nranks)

    def _set_split(self, K, k_expand, rank, nranks):
        self.K, self.Kc = K, K * k_expand
        self.rank, self.nranks = rank, nranks
        self.rows_per_rank = -(-self.M // nranks)
        self.row0 = min(self.M, rank * self.rows_per_rank)
        self.rows = min(self.M, self.row0 + self.rows_per_rank) - self.row0

    def download_knn(self):
        col = np.zeros((self.rows, self.Kc), dtype=np.int32)
        cor = np.zeros((self.rows, self.Kc), dtype=np.float32)
        check(lib().mcb_cgraph_download_knn(self.h, ptr(col, C.c_int32), ptr(cor, C.c_float)))
        return col, cor

    def knn_stats(self):
        nf, nc, cap, ns = C.c_int64(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        check(lib().mcb_cgraph_knn_stats(self.h, C.byref(nf), C.byref(nc), C.byref(cap), C.byref(ns)))
        return dict(fallback_rows=int(nf.value), candidates=int(nc.value), cap=int(cap.value), sample_cols=int(ns.value))

    def mutual(self, k_beta=3):
        check(lib().mcb_cgraph_mutual(self.h, C.c_int32(k_beta)))

    def prune(self):
        ne = C.c_int64(0)
        check(lib().mcb_cgraph_prune(self.h, C.byref(ne)))
        self.n_edges = int(ne.value)
        return self.n_edges

    def edges(self, out=None):
        """(src, dst, w) of the owned rows; `out` = optional preallocated (pinned) numpy triple"""
        if out is None:
            out = (np.zeros(self.n_edges, np.int32), np.zeros(self.n_edges, np.int32), np.zeros(self.n_edges, np.float64))
        src, dst, w = out
        check(lib().mcb_cgraph_edges(self.h, ptr(src, C.c_int32), ptr(dst, C.c_int32), ptr(w, C.c_double)))
        return src[:self.n_edges], dst[:self.n_edges], w[:self.n_edges]


def _take(pointer, n, copy=True):
    """library-allocated host array -> numpy (copied, then released with mcb_free)"""
    try:
        a = np.ctypeslib.as_array(pointer, shape=(max(n, 1),))[:n]
        return a.copy() if copy else a
    finally:
        if copy:
            lib().mcb_free(pointer)


def cgraph_bknn_oneshot(ctx, ngenes, indptr, indices, data, feat_rows, K, k_expand=10, k_beta=3, k_nonz=7.0,
                        dsamp=False, seed=42, raw=False):
    """mcb_cgraph_bknn (ctx = Context) or mcb_cgraph_bknn_multi (ctx = Multi): the single call the R glue makes, ordinary
    host arrays in, host edge table out.  raw=True skips the numpy copies and frees the library's arrays at once
    (timing the call itself): returns (n_edges, n_valid)."""
    p = as_c(indptr, np.int64)
    i = as_c(indices, np.int32)
    x = as_c(data, np.float64)
    fr = as_c(feat_rows, np.int32)
    ne, nv = C.c_int64(0), C.c_int64(0)
    ps, pd, pw = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
    fn = lib().mcb_cgraph_bknn_multi if isinstance(ctx, Multi) else lib().mcb_cgraph_bknn
    check(fn(ctx.h, C.c_int32(ngenes), C.c_int64(p.size - 1), ptr(p, C.c_int64), ptr(i, C.c_int32),
             ptr(x, C.c_double), ptr(fr, C.c_int32), C.c_int32(fr.size), C.c_int32(K),
             C.c_int32(k_expand), C.c_int32(k_beta), C.c_double(k_nonz), C.c_int(int(dsamp)),
             C.c_uint64(int(seed)), C.byref(ne), C.byref(ps), C.byref(pd), C.byref(pw), C.byref(nv)))
    n = int(ne.value)
    if raw:
        for q in (ps, pd, pw):
            lib().mcb_free(q)
        return n, int(nv.value)
    return _take(ps, n), _take(pd, n), _take(pw, n), int(nv.value)


class Multi:
    """mcb_multi: one process driving several devices (contexts + NCCL communicators + one host thread each)."""

    def __init__(self, devices):
        dev = as_c(list(devices), np.int32)
        self.h = C.c_void_p()
        check(lib().mcb_multi_create(ptr(dev, C.c_int32), C.c_int32(dev.size), C.byref(self.h)))
        self.devices = [int(d) for d in dev]

    def close(self):
        if self.h:
            lib().mcb_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(lib().mcb_multi_size(self.h))

    def ctx(self, rank):
        return Context(self.devices[rank], borrowed=lib().mcb_multi_ctx(self.h, C.c_int32(rank)))

    def set_option(self, key, value):
        check(lib().mcb_multi_set_option(self.h, key.encode(), C.c_int64(int(value))))

    def coclust(self, N, src, dst, w, samples, seeds, min_mc_size, cooling, burn_in, max_sweeps=1000, knn=0):
        """mcb_coclust_multi: (node1, node2, cnt, samples_per_node)"""
        src, dst, w = as_c(src, np.int32), as_c(dst, np.int32), as_c(w, np.float64)
        samples, seeds = as_c(samples, np.int32), as_c(seeds, np.uint64)
        n_resamp, n_s = samples.shape
        npairs = C.c_int64(0)
        p1, p2, pc, pn = (C.POINTER(C.c_int32)() for _ in range(4))
        check(lib().mcb_coclust_multi(self.h, C.c_int32(N), C.c_int64(src.size), ptr(src, C.c_int32), ptr(dst, C.c_int32),
                                      ptr(w, C.c_double), C.c_int32(n_resamp), C.c_int32(n_s), ptr(samples, C.c_int32),
                                      ptr(seeds, C.c_uint64), C.c_int32(min_mc_size), C.c_double(cooling), C.c_int32(burn_in),
                                      C.c_int32(max_sweeps), C.c_int32(knn), C.byref(npairs), C.byref(p1), C.byref(p2),
                                      C.byref(pc), C.byref(pn)))
        n = int(npairs.value)
        return _take(p1, n), _take(p2, n), _take(pc, n), _take(pn, N)


class Comm:
    """mcb_comm: this process's rank of a multi-process communicator (one process per device)."""

    def __init__(self, ctx, nranks, rank, unique_id):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        self.h = C.c_void_p()
        idb = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        check(lib().mcb_comm_create(ctx.h, C.c_int32(nranks), C.c_int32(rank), idb, C.byref(self.h)))

    @staticmethod
    def unique_id():
        b = (C.c_uint8 * 128)()
        check(lib().mcb_comm_unique_id(b))
        return bytes(b)

    def close(self):
        if self.h:
            lib().mcb_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_sharded(self, local_csc, cell_base, feat_rows, K, k_expand=10, k_beta=3, k_nonz=7.0, dsamp=False, seed=42):
        """mcb_cgraph_build_sharded: rank-collective A1..A7; returns a CGraph holding this rank's edges"""
        fr = as_c(feat_rows, np.int32)
        h = C.c_void_p()
        ne = C.c_int64(0)
        check(lib().mcb_cgraph_build_sharded(self.h, local_csc.h, C.c_int64(int(cell_base)), ptr(fr, C.c_int32), C.c_int32(fr.size),
                                             C.c_double(k_nonz), C.c_int32(K), C.c_int32(k_expand), C.c_int32(k_beta),
                                             C.c_int(int(dsamp)), C.c_uint64(int(seed)), C.byref(h), C.byref(ne)))
        g = CGraph(self.ctx, h, keep=local_csc)
        g._set_split(K, k_expand, self.rank, self.nranks)
        g.n_edges = int(ne.value)
        return g

    def gather_edges(self, g, root=0, raw=False):
        """mcb_cgraph_gather_edges: (src, dst, w) on root, None elsewhere"""
        ne = C.c_int64(0)
        ps, pd, pw = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
        check(lib().mcb_cgraph_gather_edges(self.h, g.h, C.c_int32(root), C.byref(ne), C.byref(ps), C.byref(pd), C.byref(pw)))
        n = int(ne.value)
        if self.rank != root:
            return None
        if raw:
            for q in (ps, pd, pw):
                lib().mcb_free(q)
            return n
        return _take(ps, n), _take(pd, n), _take(pw, n)

    def coclust_reduce(self, cc, root=0):
        check(lib().mcb_coclust_reduce(self.h, cc.h, C.c_int32(root)))


GENE_STAT_COLUMNS = ("tot", "var", "niche_stat", "n_mean", "is_on_count", "sz_cor", "ds_top1", "ds_top2", "ds_top3",
                     "ds_mean", "ds_var", "ds_is_on_count")


def gene_stats(ctx, csc, ds=None, niche_quantile=0.2, k_std_n=1000.0, sub_cells=None):
    """mcb_gene_stats: dict of per-gene arrays (all genes, unrounded); ds = csc.downsample(n, seed, add_non_dsamp=False)."""
    out = np.empty((len(GENE_STAT_COLUMNS), csc.ngenes), dtype=np.float64)
    sub = None if sub_cells is None else as_c(sub_cells, np.int32)
    check(lib().mcb_gene_stats(ctx.h, csc.h, ds.h if ds is not None else None, C.c_double(niche_quantile), C.c_double(k_std_n),
                               ptr(sub, C.c_int32) if sub is not None else None, C.c_int64(0 if sub is None else sub.size),
                               ptr(out, C.c_double)))
    return {k: out[q] for q, k in enumerate(GENE_STAT_COLUMNS)}


def mc_gene_tapply(ctx, csc, mc_of_cell, n_mc):
    """mcb_mc_gene_tapply: (geomean [n_mc][ngenes], frac_on [n_mc][ngenes], mc_size [n_mc], mc_meansize [n_mc])."""
    mc = as_c(mc_of_cell, np.int32)
    if mc.size != csc.ncells:
        raise McbError("MC-ERR: one metacell id per cell is required")
    geo = np.empty((int(n_mc), csc.ngenes), dtype=np.float64)
    frac = np.empty((int(n_mc), csc.ngenes), dtype=np.float64)
    size = np.empty(int(n_mc), dtype=np.int64)
    mean = np.empty(int(n_mc), dtype=np.float64)
    check(lib().mcb_mc_gene_tapply(ctx.h, csc.h, ptr(mc, C.c_int32), C.c_int32(int(n_mc)), ptr(geo, C.c_double), ptr(frac, C.c_double),
                                   ptr(size, C.c_int64), ptr(mean, C.c_double)))
    return geo, frac, size, mean


def cor_knn_xy(ctx, x_cells, y_cells, knn):
    """mcb_cor_knn_xy: x_cells [nx][G], y_cells [ny][G] (one row per cell) -> col [nx][knn] int32, cor [nx][knn] f32."""
    x = as_c(x_cells, np.float64)
    y = as_c(y_cells, np.float64)
    if x.ndim != 2 or y.ndim != 2 or x.shape[1] != y.shape[1]:
        raise McbError("MC-ERR: tgs_cor_knn: x and y must have the same number of rows (genes)")
    col = np.empty((x.shape[0], int(knn)), dtype=np.int32)
    cor = np.empty((x.shape[0], int(knn)), dtype=np.float32)
    check(lib().mcb_cor_knn_xy(ctx.h, ptr(x, C.c_double), C.c_int64(x.shape[0]), ptr(y, C.c_double), C.c_int64(y.shape[0]),
                               C.c_int32(x.shape[1]), C.c_int32(int(knn)), ptr(col, C.c_int32), ptr(cor, C.c_float)))
    return col, cor


class CoClust:
    """mcb_coclust: one resampled co-clustering run (this rank's share of the resamples)."""

    def __init__(self, ctx, N, src, dst, w, samples, seeds, min_mc_size, cooling, burn_in, max_sweeps=1000,
                 rank=0, nranks=1, knn=0, method="full"):
        src = as_c(src, np.int32)
        dst = as_c(dst, np.int32)
        w = as_c(w, np.float64)
        samples = as_c(samples, np.int32)
        seeds = as_c(seeds, np.uint64)
        n_resamp, n_s = samples.shape
        self.ctx, self.N, self.n_resamp = ctx, N, n_resamp
        self.h = C.c_void_p()
        self.method, self.n_edges = method, int(src.size)
        run = {"full": lib().mcb_coclust_run2, "edges": lib().mcb_coclust_run_edges}[method]
        check(run(ctx.h, C.c_int32(N), C.c_int64(src.size), ptr(src, C.c_int32), ptr(dst, C.c_int32),
                                     ptr(w, C.c_double), C.c_int32(n_resamp), C.c_int32(n_s), ptr(samples, C.c_int32),
                                     ptr(seeds, C.c_uint64), C.c_int32(min_mc_size), C.c_double(cooling),
                                     C.c_int32(burn_in), C.c_int32(max_sweeps), C.c_int32(knn), C.c_int32(rank), C.c_int32(nranks),
                                     C.byref(self.h)))
        self.n_local = len(range(rank, n_resamp, nranks))

    def destroy(self):
        if self.h:
            lib().mcb_coclust_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def buffers(self):
        a, b = C.c_void_p(), C.c_void_p()
        check(lib().mcb_coclust_buffers(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def assignments(self):
        mc = np.zeros((max(self.n_local, 1), self.N), dtype=np.int32)
        sw = np.zeros(max(self.n_local, 1), dtype=np.int32)
        nl = C.c_int32(0)
        check(lib().mcb_coclust_assignments(self.h, ptr(mc, C.c_int32), ptr(sw, C.c_int32), C.byref(nl)))
        return mc[:nl.value], sw[:nl.value]

    def edge_counts(self):
        """method = "edges": (node1, node2, cnt) for every edge of the graph (cnt may be 0) + samples[N]"""
        n1, n2, cnt = (np.zeros(max(self.n_edges, 1), np.int32) for _ in range(3))
        samples = np.zeros(self.N, np.int32)
        check(lib().mcb_coclust_edge_counts(self.h, ptr(n1, C.c_int32), ptr(n2, C.c_int32), ptr(cnt, C.c_int32), ptr(samples, C.c_int32)))
        n = self.n_edges
        return n1[:n], n2[:n], cnt[:n], samples

    def pairs(self):
        npairs = C.c_int64(0)
        check(lib().mcb_coclust_extract(self.h, C.byref(npairs)))
        n = int(npairs.value)
        n1, n2, cnt = (np.zeros(max(n, 1), np.int32) for _ in range(3))
        samples = np.zeros(self.N, np.int32)
        check(lib().mcb_coclust_pairs(self.h, ptr(n1, C.c_int32), ptr(n2, C.c_int32), ptr(cnt, C.c_int32),
                                      ptr(samples, C.c_int32)))
        return n1[:n], n2[:n], cnt[:n], samples


def graph_cover(ctx, N, src, dst, w, min_mc_size, cooling, burn_in, seed, max_sweeps=1000):
    """mcb_graph_cover: one cover over all nodes; returns (cluster[N] with 0 = outlier, sweeps)."""
    src = as_c(src, np.int32)
    dst = as_c(dst, np.int32)
    w = as_c(w, np.float64)
    out = np.zeros(N, dtype=np.int32)
    sw = C.c_int32(0)
    check(lib().mcb_graph_cover(ctx.h, C.c_int32(N), C.c_int64(src.size), ptr(src, C.c_int32), ptr(dst, C.c_int32),
                                ptr(w, C.c_double), C.c_int32(min_mc_size), C.c_double(cooling), C.c_int32(burn_in),
                                C.c_int32(max_sweeps), C.c_uint64(int(seed)), ptr(out, C.c_int32), C.byref(sw)))
    return out, int(sw.value)
