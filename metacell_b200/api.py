"""Host-side mirror of the reference's R interface for the bknn-cgraph + co-clustering path.

The reference is an R package; R is not available in this image, so the host layer that R would
provide is written here in Python with the reference's names, argument meaning and error behaviour
(`stop("MC-ERR: ...")` -> McbError("MC-ERR: ...")).  The R `.Call` glue a maintainer would add is in
r/ and INTEGRATION.md.  All arithmetic goes through libmcb200.so (include/mcb200.h); nothing here
falls back to numpy or the oracle.

Mirrored entry points (reference file:line):
    scdb_init / scdb_add_* / scdb_*            R/scdb.r:100-131, 292-320, 400-428
    get_param / set_param                      inst/config/metacell_params.yaml (keys of the path)
    tgScMat, tgGeneSets, tgCellGraph, tgCoClust  R/scmat.r, R/gset.r, R/cgraph.r:11-42, R/coclust.r:12-48
    scm_which_downsamp_n                       R/scmat.r:399-407
    scm_downsamp                               R/utils.r:30-69
    gset_get_feat_mat                          R/gset.r:151-214
    tgs_cor_knn / tgs_graph / tgs_cor_graph    tgstat calls at R/utils.r:115-123
    mcell_add_cgraph_from_mat_bknn             R/cgraph_knn.r:10-25
    tgs_graph_cover_resample                   tgstat call at R/coclust_graph_cov.r:41
    mcell_coclust_from_graph_resamp            R/coclust_graph_cov.r:13-58
and, as the first callers on either side of the path:
    tgs_graph_cover                            tgstat call at R/mc_graphcov.r:27, R/mc_from_coclust.r:247
    mcell_add_mc_from_graph                    R/mc_graphcov.r:9-41 (up to the cluster assignment)
    mcell_coclust_filt_by_k_deg                R/coclust.r:58-84
    mcell_mc_from_coclust_balanced             R/mc_from_coclust.r:215-263 (up to the cluster assignment)
"""
import os
import pickle
import sys

import numpy as np

from . import _lib
from ._lib import McbError

# ------------------------------------------------------------------------------------------------
# parameters (inst/config/metacell_params.yaml:12,20,41,44,49-51,53-54)
# ------------------------------------------------------------------------------------------------
_PARAMS = {
    "mc_cores": 16,
    "mc_rseed": 42,
    "scm_n_downsamp_gstat": None,
    "scm_k_nonz_exp": 7,
    "scm_balance_graph_k_alpha": 10,
    "scm_balance_graph_k_beta": 3,
    "scm_balance_graph_k_expand": 10,
    "scm_tgs_clust_cool": 1.05,
    "scm_tgs_clust_burn_in": 10,
}


def get_param(key):
    if key not in _PARAMS:
        raise McbError(f"MC-ERR: unknown parameter {key}")
    return _PARAMS[key]


def set_param(key, value):
    if key not in _PARAMS:
        raise McbError(f"MC-ERR: unknown parameter {key}")
    _PARAMS[key] = value


def _message(msg):
    print(msg, file=sys.stderr)


# ------------------------------------------------------------------------------------------------
# containers
# ------------------------------------------------------------------------------------------------
class tgScMat:
    """Sparse UMI matrix, genes x cells, stored as the dgCMatrix slots (reference R/scmat.r)."""

    def __init__(self, indptr, indices, data, genes=None, cells=None, ngenes=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.ncells = self.indptr.size - 1
        self.ngenes = int(ngenes if ngenes is not None else (len(genes) if genes is not None else self.indices.max() + 1))
        self.genes = list(genes) if genes is not None else [f"g{i}" for i in range(self.ngenes)]
        self.cells = list(cells) if cells is not None else [f"c{i}" for i in range(self.ncells)]


class tgGeneSets:
    """Named gene set; `gene_set` maps gene name -> set id, in insertion order (reference R/gset.r)."""

    def __init__(self, gene_set, description=""):
        self.gene_set = dict(gene_set)
        self.description = description


class tgCellGraph:
    """Cell similarity graph (reference R/cgraph.r:11-42): `edges` holds mc1, mc2 (integer codes into
    cell_names, the R factor codes minus one) and w; `nodes` are the cells that carry an edge."""

    def __init__(self, edges, cell_names):
        for k in ("mc1", "mc2", "w"):
            if k not in edges:
                raise McbError("MC-ERR: cell graph edges need columns mc1, mc2, w")
        self.cell_names = list(cell_names)
        self.edges = {k: np.asarray(edges[k]) for k in ("mc1", "mc2", "w")}
        nodes = np.unique(np.concatenate([self.edges["mc1"], self.edges["mc2"]])) if len(self.edges["mc1"]) else np.zeros(0, np.int64)
        self.nodes = [self.cell_names[i] for i in nodes]
        if len(self.cell_names) != len(self.nodes):       # R/cgraph.r:36-39
            _message(f"sim graph is missing {len(self.cell_names) - len(self.nodes)} nodes, out of {len(self.cell_names)}")


class tgCoClust:
    """Co-clustering statistics (reference R/coclust.r:12-48)."""

    def __init__(self, graph_id, coclust, n_samp):
        if scdb_cgraph(graph_id) is None:
            raise McbError(f"MC-ERR: graph id {graph_id} is missing when constructing a coclust object")
        for k in ("node1", "node2", "cnt"):
            if k not in coclust:
                raise McbError("MC-ERR: coclust table must have columns node1, node2, cnt")
        self.graph_id = graph_id
        self.coclust = {k: np.asarray(coclust[k]) for k in ("node1", "node2", "cnt")}
        self.n_samp = np.asarray(n_samp)


# ------------------------------------------------------------------------------------------------
# scdb (reference R/scdb.r): in-memory cache + one file per object, "<base>/<type>.<id>.Rda" there,
# "<base>/<type>.<id>.pkl" here
# ------------------------------------------------------------------------------------------------
_SCDB = {"base": None, "cache": {}}


def scdb_init(base_dir, force_reinit=False):
    if not os.path.isdir(base_dir):
        raise McbError(f"MC-ERR: cannot initialize scdb to non existing directory {base_dir}")
    if _SCDB["base"] is not None and not force_reinit and _SCDB["base"] != base_dir:
        raise McbError("MC-ERR: scdb already initialized, use force_reinit to switch directory")
    _SCDB["base"] = base_dir
    _SCDB["cache"] = {}


def _scdb_fn(objt, oid):
    return os.path.join(_SCDB["base"], f"{objt}.{oid}.pkl")


def _scdb_add(objt, oid, obj):
    _SCDB["cache"][(objt, oid)] = obj
    if _SCDB["base"] is not None:
        with open(_scdb_fn(objt, oid), "wb") as fh:
            pickle.dump(obj, fh, protocol=4)


def _scdb_get(objt, oid):
    if (objt, oid) in _SCDB["cache"]:
        return _SCDB["cache"][(objt, oid)]
    if _SCDB["base"] is not None and os.path.exists(_scdb_fn(objt, oid)):
        with open(_scdb_fn(objt, oid), "rb") as fh:
            obj = pickle.load(fh)
        _SCDB["cache"][(objt, oid)] = obj
        return obj
    return None


def scdb_add_mat(oid, mat): _scdb_add("mat", oid, mat)
def scdb_mat(oid): return _scdb_get("mat", oid)
def scdb_add_gset(oid, gset): _scdb_add("gset", oid, gset)
def scdb_gset(oid): return _scdb_get("gset", oid)
def scdb_add_cgraph(oid, cgraph): _scdb_add("cgraph", oid, cgraph)
def scdb_cgraph(oid): return _scdb_get("cgraph", oid)
def scdb_add_coclust(oid, coc): _scdb_add("coclust", oid, coc)
def scdb_coclust(oid): return _scdb_get("coclust", oid)


# ------------------------------------------------------------------------------------------------
# A1 / A2 / A3
# ------------------------------------------------------------------------------------------------
def _ctx(ctx=None):
    return ctx if ctx is not None else _lib.default_context()


def _upload(mat, ctx):
    return _lib.Csc.upload(ctx, mat.ngenes, mat.indptr, mat.indices, mat.data)


def scm_which_downsamp_n(mat, ctx=None):
    """reference R/scmat.r:399-407"""
    n = get_param("scm_n_downsamp_gstat")
    if n is not None:
        return n
    ctx = _ctx(ctx)
    csc = _upload(mat, ctx)
    tot = csc.col_sums()
    csc.free()
    return _lib.downsamp_depth(tot)


def scm_downsamp(umis, n, ctx=None):
    """reference R/utils.r:30-69: returns the down-sampled tgScMat holding only the cells with at least
    n UMIs (column sums == floor(n))."""
    ctx = _ctx(ctx)
    csc = _upload(umis, ctx)
    ds = csc.downsample(float(n), get_param("mc_rseed"), add_non_dsamp=False)
    x, kept = ds.download()
    ds.free()
    csc.free()
    return _subset_cells(umis, x, np.nonzero(kept)[0])


def _subset_cells(mat, data, cols):
    lens = np.diff(mat.indptr)
    keep_entry = np.repeat(np.isin(np.arange(mat.ncells), cols), lens) & (data > 0)
    col_of_entry = np.repeat(np.arange(mat.ncells), lens)[keep_entry]
    new_lens = np.bincount(col_of_entry, minlength=mat.ncells)[cols]
    indptr = np.concatenate([[0], np.cumsum(new_lens)])
    return tgScMat(indptr, mat.indices[keep_entry], data[keep_entry], genes=mat.genes,
                   cells=[mat.cells[c] for c in cols], ngenes=mat.ngenes)


def _feature_rows(gset, mat):
    """intersect(names(gset@gene_set), rownames(umis)), gene-set order (reference R/gset.r:192-197)"""
    row_of = {g: i for i, g in enumerate(mat.genes)}
    rows = [row_of[g] for g in gset.gene_set.keys() if g in row_of]
    if len(rows) < 2:
        raise McbError("MC-ERR: less than 2 feature genes of the gene set are present in the matrix")
    return np.asarray(rows, dtype=np.int32)


def gset_get_feat_mat(gset_id, mat_id, downsamp=False, add_non_dsamp=False, downsample_n=None, ctx=None):
    """reference R/gset.r:151-214; returns the host feature matrix (tgScMat restricted to feature rows)."""
    gset, mat = scdb_gset(gset_id), scdb_mat(mat_id)
    if gset is None:
        raise McbError(f"MC-ERR: non existing gset {gset_id} when trying to get feature matrix")
    if mat is None:
        raise McbError(f"MC-ERR: non existing mat {mat_id} when trying to get feature matrix")
    ctx = _ctx(ctx)
    data, cols = mat.data, np.arange(mat.ncells)
    if downsamp:
        n = downsample_n if downsample_n is not None else scm_which_downsamp_n(mat, ctx)
        _message(f"will downsample the matrix, N= {n}")
        csc = _upload(mat, ctx)
        ds = csc.downsample(float(n), get_param("mc_rseed"), add_non_dsamp=bool(add_non_dsamp))
        data, kept = ds.download()
        ds.free()
        csc.free()
        if not add_non_dsamp:
            cols = np.nonzero(kept)[0]
    sub = _subset_cells(mat, data, cols)
    rows = _feature_rows(gset, mat)
    # row subset on the host: a container operation, not path arithmetic
    pos = np.full(mat.ngenes, -1, dtype=np.int64)
    pos[rows] = np.arange(rows.size)
    keep = pos[sub.indices] >= 0
    col_of_entry = np.repeat(np.arange(sub.ncells), np.diff(sub.indptr))[keep]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(col_of_entry, minlength=sub.ncells))])
    return tgScMat(indptr, pos[sub.indices[keep]].astype(np.int32), sub.data[keep],
                   genes=[mat.genes[r] for r in rows], cells=sub.cells, ngenes=rows.size)


# ------------------------------------------------------------------------------------------------
# mc_compute_fp / mc_compute_e_gc / mc_compute_cov_gc / mc_update_stats (reference R/mc.r:92-241)
# ------------------------------------------------------------------------------------------------
def _mc_tapply(mc, mat, ctx):
    """one GPU pass shared by the three summaries: us = mat[, names(mc@mc)] (R/mc.r:94) is expressed as a per-cell
    metacell id with -1 for cells outside the cover"""
    index = {nm: q for q, nm in enumerate(mat.cells)}
    ids = np.full(mat.ncells, -1, dtype=np.int32)
    for nm, k in mc["mc"].items():
        q = index.get(nm)
        if q is None:
            raise McbError(f"MC-ERR: cell {nm} of the metacell cover is missing from the matrix")
        ids[q] = int(k) - 1
    n_mc = int(ids.max()) + 1
    csc = _upload(mat, ctx)
    try:
        geo, frac, size, meansize = _lib.mc_gene_tapply(ctx, csc, ids, n_mc)
        covered = ids >= 0
        # rowSums(us): gene totals over the covered cells only
        tot = np.bincount(mat.indices[np.repeat(covered, np.diff(mat.indptr))],
                          weights=mat.data[np.repeat(covered, np.diff(mat.indptr))], minlength=mat.ngenes)
    finally:
        csc.free()
    return geo, frac, size, meansize, tot


def mc_compute_fp(mc, mat, norm_by_mc_size=True, min_total_umi=10, ctx=None, _pre=None):
    """reference R/mc.r:164-207: regularised, per-gene median-normalised geometric-mean footprint, genes x metacells."""
    geo, _, _, meansize, tot = _pre if _pre is not None else _mc_tapply(mc, mat, _ctx(ctx))
    f_g = tot > min_total_umi
    g = geo[:, f_g].T                                                     # genes x metacells (t(tgs_matrix_tapply(...)))
    if norm_by_mc_size:
        ideal = min(1000.0, float(np.median(meansize)))
        g = ideal * g / meansize[None, :]
    fp_reg = 0.1
    g = (fp_reg + g) / np.median(fp_reg + g, axis=1, keepdims=True)
    return g, [mat.genes[i] for i in np.nonzero(f_g)[0]]


def mc_compute_e_gc(mc, mat, norm_by_mc_meansize=True, ctx=None, _pre=None):
    """reference R/mc.r:218-229"""
    geo, _, _, meansize, tot = _pre if _pre is not None else _mc_tapply(mc, mat, _ctx(ctx))
    f_g = tot > 10
    e = geo[:, f_g].T
    if norm_by_mc_meansize:
        e = e / meansize[None, :]
    return e, [mat.genes[i] for i in np.nonzero(f_g)[0]]


def mc_compute_cov_gc(mc, mat, ctx=None, _pre=None):
    """reference R/mc.r:237-243"""
    _, frac, _, _, tot = _pre if _pre is not None else _mc_tapply(mc, mat, _ctx(ctx))
    f_g = tot > 10
    return frac[:, f_g].T, [mat.genes[i] for i in np.nonzero(f_g)[0]]


def mc_update_stats(mc, mat, ctx=None):
    """reference R/mc.r:92-105 (batch counts excluded): footprints, absolute expression and coverage from one pass"""
    pre = _mc_tapply(mc, mat, _ctx(ctx))
    mc = dict(mc)
    mc["mc_fp"], mc["genes"] = mc_compute_fp(mc, mat, _pre=pre)
    mc["e_gc"], _ = mc_compute_e_gc(mc, mat, _pre=pre)
    mc["cov_gc"], _ = mc_compute_cov_gc(mc, mat, _pre=pre)
    return mc


# ------------------------------------------------------------------------------------------------
# scm_gene_stat (reference R/gstats.r:65-205)
# ------------------------------------------------------------------------------------------------
def _rollmedian(x, k, left, right):
    """zoo::rollmedian(x, k, fill = c(left, NA, right)), k odd (R/gstats.r:157,168,182)"""
    h = k // 2
    out = np.empty(x.size, dtype=np.float64)
    out[:h] = left
    out[x.size - h:] = right
    if x.size > 2 * h:
        out[h:x.size - h] = np.median(np.lib.stride_tricks.sliding_window_view(x, k), axis=1)
    return out


def _trend_norm(values, order):
    """values[order] - rolling median (window 101) of values[order], ends filled with the medians of the first 101 /
    last 102 entries exactly as R/gstats.r:153-161 index them."""
    v = values[order]
    cmin = np.median(v[:101])
    cmax = np.median(v[v.size - 102:])
    out = np.empty_like(values)
    out[order] = v - _rollmedian(v, 101, cmin, cmax)
    return out


def scm_gene_stat(mat_id, niche_quantile=0.2, downsample_n=None, K_std_n=1000, ctx=None):
    """reference R/gstats.r:65-205.  Returns a dict of equal-length columns (the R data frame), one entry per gene with
    more than 10 UMIs (:95).  The per-gene sums, the downsampling and the niche selection run on the GPU
    (mcb_gene_stats); the O(genes) rolling-median trends are host work as in the reference.  Like the reference the
    `downsample_n` argument is overwritten by scm_which_downsamp_n (:97)."""
    mat = scdb_mat(mat_id)
    if mat is None:
        raise McbError(f"MC-ERR: missing mat {mat_id} when computing gene statistics")
    if niche_quantile > 1 or niche_quantile < 0:
        raise McbError(f"niche_quantile must be between 0 and 1, got {niche_quantile}")
    ctx = _ctx(ctx)
    _message("Calculating gene statistics... ")
    csc = _upload(mat, ctx)
    ds = None
    try:
        csum = csc.col_sums()
        f_oversize = (csum > np.quantile(csum, 0.95) * 2) | (csum < 100)                 # :82-83
        N = int((~f_oversize).sum())
        downsample_n = _lib.downsamp_depth(csum) if get_param("scm_n_downsamp_gstat") is None else get_param("scm_n_downsamp_gstat")
        _message("will downsamp")
        ds = csc.downsample(float(downsample_n), get_param("mc_rseed"), add_non_dsamp=False)
        _message("done downsamp")
        sub = None
        if mat.ncells > 50000:                                                            # :143-146 (host-supplied index set)
            sub = np.sort(np.random.default_rng(int(get_param("mc_rseed"))).choice(mat.ncells, 50000, replace=False)).astype(np.int32)
        st = _lib.gene_stats(ctx, csc, ds, float(niche_quantile), float(K_std_n), sub)
    finally:
        if ds is not None:
            ds.free()
        csc.free()
    f_g = st["tot"] > 10                                                                  # :95
    if int(f_g.sum()) < 102:
        raise McbError("MC-ERR: fewer than 102 genes with more than 10 UMIs, cannot compute gene statistic trends")
    g = {k: v[f_g] for k, v in st.items()}
    res = {"name": [mat.genes[i] for i in np.nonzero(f_g)[0]]}
    res["tot"] = g["tot"]
    res["var"] = np.round(g["var"], 7)
    res["is_on_count"] = g["is_on_count"]
    res["sz_cor"] = np.round(g["sz_cor"], 3)
    tot_ord = np.argsort(res["tot"], kind="stable")
    res["sz_cor_norm"] = _trend_norm(res["sz_cor"], tot_ord)
    res["niche_stat"] = g["niche_stat"]
    res["niche_norm"] = _trend_norm(res["niche_stat"], tot_ord)
    res["n_mean"] = np.round(g["n_mean"], 7)
    for k in ("ds_top1", "ds_top2", "ds_top3"):
        res[k] = g[k]
    res["ds_mean"] = np.round(g["ds_mean"], 7)
    res["ds_var"] = np.round(g["ds_var"], 7)
    m_reg = 10.0 / N                                                                      # :176
    res["ds_log_varmean"] = np.log2((m_reg + res["ds_var"]) / (m_reg + res["ds_mean"]))
    ds_ord = np.argsort(res["ds_mean"], kind="stable")
    res["ds_vm_norm"] = _trend_norm(res["ds_log_varmean"], ds_ord)
    res["ds_is_on_count"] = g["ds_is_on_count"]
    res["downsample_n"] = np.full(res["tot"].size, float(downsample_n))
    _message("..done")
    return res


def mcell_add_gene_stat(mat_id, gstat_id, force=False, ctx=None):
    """reference R/gstats.r:19-27: compute and store the gene statistics table in scdb."""
    if not force and _scdb_get("gstat", gstat_id) is not None:
        return _scdb_get("gstat", gstat_id)
    gs = scm_gene_stat(mat_id, ctx=ctx)
    _scdb_add("gstat", gstat_id, gs)
    return gs


# ------------------------------------------------------------------------------------------------
# A5 / A6: the tgstat trio
# ------------------------------------------------------------------------------------------------
def tgs_cor_knn(x, y=None, knn=None, ctx=None):
    """tgs_cor_knn(x, y = x, knn): x (and y) are genes x cells (like the R matrices).  Returns dict col1, col2, cor,
    rank (schema pinned by reference R/atlas.r:282,477,585); col1/col2 are 0-based cell codes (col2 indexes y's
    columns).  y is x: self is rank 1 (R/atlas.r:475-476).  y is not x (atlas projection, R/atlas.r:164-166): the knn
    best y cells per x cell, no forced self; NA correlations (constant cells) produce no rows."""
    if y is not None and y is not x:
        ctx = _ctx(ctx)
        xa, ya = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if xa.shape[0] != ya.shape[0]:
            raise McbError("MC-ERR: tgs_cor_knn: x and y must have the same number of rows (genes)")
        col, cor = _lib.cor_knn_xy(ctx, np.ascontiguousarray(xa.T), np.ascontiguousarray(ya.T), int(knn))
        ok = col >= 0
        col1 = np.broadcast_to(np.arange(col.shape[0], dtype=np.int32)[:, None], col.shape)
        rank = np.broadcast_to(np.arange(1, col.shape[1] + 1, dtype=np.int32)[None, :], col.shape)
        return {"col1": col1[ok].astype(np.int32), "col2": col[ok], "cor": cor[ok].astype(np.float64), "rank": rank[ok].astype(np.int32)}
    ctx = _ctx(ctx)
    xt = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
    g = _lib.CGraph.from_dense(ctx, xt)
    try:
        if g.M < 2:
            raise McbError("MC-ERR: fewer than 2 cells with non-constant features")
        g.knn(int(knn), 1)
        col, cor = g.download_knn()
        cells = g.valid_cells()
    finally:
        g.destroy()
    kk = min(int(knn), g.M)
    col1 = np.repeat(cells, kk)
    col2 = cells[col[:, :kk].reshape(-1)]
    rank = np.tile(np.arange(1, kk + 1, dtype=np.int32), g.M)
    return {"col1": col1.astype(np.int32), "col2": col2.astype(np.int32), "cor": cor[:, :kk].reshape(-1).astype(np.float64),
            "rank": rank}


def _graph_from_handle(g, K, k_expand, k_beta):
    g.knn(int(K), int(k_expand))
    g.mutual(int(k_beta))
    g.prune()
    src, dst, w = g.edges()
    return {"col1": src.copy(), "col2": dst.copy(), "weight": w.copy()}


def tgs_cor_graph(x, knn, k_expand, k_alpha, k_beta, ctx=None):
    """reference R/utils.r:115-123: tgs_cor_knn(x, x, knn*k_expand) then tgs_graph(...); k_alpha is
    accepted and ignored exactly as the reference wrapper does.  x: genes x cells dense."""
    ctx = _ctx(ctx)
    xt = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
    g = _lib.CGraph.from_dense(ctx, xt)
    try:
        return _graph_from_handle(g, knn, k_expand, k_beta)
    finally:
        g.destroy()


# ------------------------------------------------------------------------------------------------
# mcell_add_cgraph_from_mat_raw_knn (reference R/cgraph_knn.r:35-51)
# ------------------------------------------------------------------------------------------------
def mcell_add_cgraph_from_mat_raw_knn(mat_id, gset_id, graph_id, K, dsamp=False, ctx=None):
    """Unbalanced K-nn graph: tgs_knn(tgs_cor(feat), K) keeps, per cell, the K best correlations of the full matrix
    (the diagonal included, so self is rank 1) and weights them 1 - (rank-1)/K (R/cgraph_knn.r:43-46)."""
    gset, mat = scdb_gset(gset_id), scdb_mat(mat_id)
    if gset is None:
        raise McbError(f"MC-ERR: non existing gset ({gset_id}) when trying to build a cell graph")
    if mat is None:
        raise McbError(f"MC-ERR: non existing mat ({mat_id}) when trying to build a cell graph")
    ctx = _ctx(ctx)
    rows = _feature_rows(gset, mat)
    _message(f"will build raw knn graph on {mat.ncells} cells and {rows.size} genes, this can be a bit heavy for >20,000 cells")
    csc = _upload(mat, ctx)
    use = csc
    if dsamp:       # gset_get_feat_mat(downsamp = dsamp, add_non_dsamp = dsamp), R/cgraph_knn.r:37
        n = get_param("scm_n_downsamp_gstat")
        if n is None:
            n = _lib.downsamp_depth(csc.col_sums())
        use = csc.downsample(float(n), get_param("mc_rseed"), add_non_dsamp=True)
    g = _lib.CGraph.from_csc(ctx, use, rows, float(get_param("scm_k_nonz_exp")))
    try:
        g.knn(int(K), 1)
        col, _ = g.download_knn()
        cells = g.valid_cells()
    finally:
        g.destroy()
        if use is not csc:
            use.free()
        csc.free()
    kk = min(int(K), cells.size)
    src = np.repeat(cells, kk)
    dst = cells[col[:, :kk].reshape(-1)]
    w = np.tile(1.0 - np.arange(kk, dtype=np.float64) / float(K), cells.size)       # gr$w = 1-(rank-1)/K
    gr = tgCellGraph({"mc1": src.astype(np.int32), "mc2": dst.astype(np.int32), "w": w}, mat.cells)
    scdb_add_cgraph(graph_id, gr)
    return gr


# ------------------------------------------------------------------------------------------------
# mcell_mc_confusion_mat (reference R/mc_confusion.r:10-45)
# ------------------------------------------------------------------------------------------------
def mcell_mc_confusion_mat(mc_id, graph_id, K, ignore_mismatch=False):
    """confu[a, b] = number of graph edges from a cell of metacell a to a cell of metacell b (1-based metacell ids
    become 0-based rows here).  Host tabulation, as in the reference; the reference computes the edge filter
    `w > 1-(K+1)/deg` and then overwrites it (R/mc_confusion.r:27-37), so all edges between covered cells count."""
    graph = scdb_cgraph(graph_id)
    if graph is None:
        raise McbError(f"MC-ERR: cell graph id {graph_id} is missing when computing mc confusion")
    mc = _scdb_get("mc", mc_id)
    if mc is None:
        raise McbError(f"MC-ERR: mc id {mc_id} is missing when computing mc confusion")
    names = graph.cell_names
    mc_map = np.full(len(names), -1, dtype=np.int64)
    index = {nm: q for q, nm in enumerate(names)}
    missing = 0
    for nm, k in mc["mc"].items():
        q = index.get(nm)
        if q is None:
            missing += 1
        else:
            mc_map[q] = k
    if not ignore_mismatch and missing:
        raise McbError("MC-ERR: graph is missing some of the cells in the mc cover when computing confusion, halting")
    m1, m2 = mc_map[graph.edges["mc1"]], mc_map[graph.edges["mc2"]]
    ok = (m1 > 0) & (m2 > 0)
    N = int(max(m1[ok].max(initial=0), m2[ok].max(initial=0)))
    confu = np.bincount((m1[ok] - 1) * N + (m2[ok] - 1), minlength=N * N).reshape(N, N)
    return confu


# ------------------------------------------------------------------------------------------------
# mcell_add_cgraph_from_mat_bknn (reference R/cgraph_knn.r:10-25)
# ------------------------------------------------------------------------------------------------
def mcell_add_cgraph_from_mat_bknn(mat_id, gset_id, graph_id, K, dsamp=False, ctx=None):
    gset, mat = scdb_gset(gset_id), scdb_mat(mat_id)
    if gset is None:
        raise McbError(f"MC-ERR: non existing gset {gset_id} when trying to get feature matrix")
    if mat is None:
        raise McbError(f"MC-ERR: non existing mat {mat_id} when trying to get feature matrix")
    ctx = _ctx(ctx)
    k_nonz_exp = get_param("scm_k_nonz_exp")
    k_expand = get_param("scm_balance_graph_k_expand")
    k_beta = get_param("scm_balance_graph_k_beta")
    rows = _feature_rows(gset, mat)
    _message(f"will build balanced knn graph on {mat.ncells} cells and {rows.size} genes, this can be a bit heavy for >20,000 cells")
    csc = _upload(mat, ctx)
    use = csc
    if dsamp:       # gset_get_feat_mat(downsamp = dsamp, add_non_dsamp = dsamp), R/cgraph_knn.r:12
        n = get_param("scm_n_downsamp_gstat")
        if n is None:
            n = _lib.downsamp_depth(csc.col_sums())
        _message(f"will downsample the matrix, N= {n}")
        use = csc.downsample(float(n), get_param("mc_rseed"), add_non_dsamp=True)
    g = _lib.CGraph.from_csc(ctx, use, rows, float(k_nonz_exp))
    try:
        gr = _graph_from_handle(g, K, k_expand, k_beta)
    finally:
        g.destroy()
        if use is not csc:
            use.free()
        csc.free()
    edges = {"mc1": gr["col1"], "mc2": gr["col2"], "w": gr["weight"]}       # colnames(gr) = c("mc1","mc2","w")
    cg = tgCellGraph(edges, mat.cells)
    scdb_add_cgraph(graph_id, cg)
    return cg


# ------------------------------------------------------------------------------------------------
# A8: co-clustering
# ------------------------------------------------------------------------------------------------
def resample_index_sets(ncells, p_resamp, n_resamp, seed):
    """Host-supplied node samples + per-resample seeds (north_star: 'host-supplied RNG seeds and resample
    index sets').  Sample size round(p * N); each set strictly ascending."""
    rng = np.random.default_rng(seed)
    n_s = int(round(p_resamp * ncells))
    sets = np.empty((n_resamp, n_s), dtype=np.int32)
    for r in range(n_resamp):
        sets[r] = np.sort(rng.choice(ncells, size=n_s, replace=False))
    seeds = rng.integers(0, 2 ** 63 - 1, size=n_resamp, dtype=np.int64).astype(np.uint64)
    return sets, seeds


def tgs_graph_cover_resample(edges, knn, min_cluster_size, cooling, burn_in, p_resamp, n_resamp, method="full",
                             n_nodes=None, seed=None, ctx=None, max_sweeps=1000):
    """tgstat call of reference R/coclust_graph_cov.r:41.  edges: dict col1, col2, weight (0-based codes).
    `knn` (R passes K = round(nrow(edges) / ncells), the average out-degree, R/coclust_graph_cov.r:39) caps the out edges a
    node may use inside one resample: its knn heaviest edges among those whose target was sampled too; 0 = no cap.  Returns dict(co_cluster = dict(node1, node2, cnt), samples = int[N])."""
    if method not in ("full", "edges"):
        raise McbError("MC-ERR: method must be \"full\" (what metacell calls) or \"edges\" (co-occurrence of graph edges only)")
    ctx = _ctx(ctx)
    src, dst, w = edges["col1"], edges["col2"], edges["weight"]
    N = int(n_nodes if n_nodes is not None else max(int(np.max(src)), int(np.max(dst))) + 1)
    sets, seeds = resample_index_sets(N, p_resamp, n_resamp, get_param("mc_rseed") if seed is None else seed)
    cc = _lib.CoClust(ctx, N, src, dst, w, sets, seeds, int(min_cluster_size), float(cooling), int(burn_in), max_sweeps,
                      knn=int(knn or 0), method=method)
    try:
        if method == "full":
            n1, n2, cnt, samples = cc.pairs()
        else:                                   # O(edges) memory: the variant for cell counts whose N x N matrix cannot exist
            n1, n2, cnt, samples = cc.edge_counts()
            keep = cnt > 0
            n1, n2, cnt = n1[keep], n2[keep], cnt[keep]
    finally:
        cc.destroy()
    return {"co_cluster": {"node1": n1.copy(), "node2": n2.copy(), "cnt": cnt.copy()}, "samples": samples}


def mcell_coclust_from_graph_resamp(coc_id, graph_id, min_mc_size, p_resamp, n_resamp, ctx=None):
    """reference R/coclust_graph_cov.r:13-58"""
    cooling = get_param("scm_tgs_clust_cool")
    burn_in = get_param("scm_tgs_clust_burn_in")
    graph = scdb_cgraph(graph_id)
    if graph is None:
        raise McbError(f"MC-ERR: cell graph id {graph_id} is missing when running add_mc_from_graph")
    _message("running bootstrap to generate cocluster")
    edges = {"col1": graph.edges["mc1"], "col2": graph.edges["mc2"], "weight": graph.edges["w"]}
    ncells = len(graph.cell_names)
    K = int(round(len(edges["col1"]) / ncells))                           # R/coclust_graph_cov.r:39
    resamp = tgs_graph_cover_resample(edges, K, min_mc_size, cooling, burn_in, p_resamp, n_resamp, method="full",
                                      n_nodes=ncells, ctx=ctx)
    _message("done resampling")
    coc = tgCoClust(graph_id, resamp["co_cluster"], resamp["samples"])
    scdb_add_coclust(coc_id, coc)
    return coc


# ------------------------------------------------------------------------------------------------
# SURVEY 8f, first "next" rows: the single graph cover and the callers immediately after the path.
# Building the tgMCCov object (footprints etc., R/mc.r) is out of scope: these return the assignment.
# ------------------------------------------------------------------------------------------------
def tgs_graph_cover(edges, min_mc_size, cooling, burn_in, n_nodes=None, seed=None, ctx=None, max_sweeps=1000):
    """tgstat call of reference R/mc_graphcov.r:27: edges = dict col1, col2, weight (0-based codes).
    Returns dict(node, cluster) with cluster 0 = outlier (R/mc_graphcov.r:28)."""
    ctx = _ctx(ctx)
    src, dst, w = edges["col1"], edges["col2"], edges["weight"]
    N = int(n_nodes if n_nodes is not None else max(int(np.max(src)), int(np.max(dst))) + 1)
    cluster, _ = _lib.graph_cover(ctx, N, src, dst, w, int(min_mc_size), float(cooling), int(burn_in),
                                  get_param("mc_rseed") if seed is None else seed, max_sweeps)
    return {"node": np.arange(N, dtype=np.int32), "cluster": cluster}


def _mc_from_cover(node_clust, cell_names):
    f_outlier = node_clust["cluster"] == 0
    outliers = [cell_names[i] for i in node_clust["node"][f_outlier]]
    _, mc = np.unique(node_clust["cluster"][~f_outlier], return_inverse=True)          # as.integer(as.factor(cluster))
    names = [cell_names[i] for i in node_clust["node"][~f_outlier]]
    return {"mc": dict(zip(names, (mc + 1).tolist())), "outliers": outliers}


def mcell_add_mc_from_graph(mc_id, graph_id, mat_id, min_mc_size, ctx=None):
    """reference R/mc_graphcov.r:9-41 up to the cluster assignment (the tgMCCov object itself is out of scope)."""
    graph = scdb_cgraph(graph_id)
    if graph is None:
        raise McbError(f"MC-ERR: cell graph id {graph_id} is missing when running add_mc_from_graph")
    mat = scdb_mat(mat_id)
    if mat is None:
        raise McbError(f"MC-ERR: mat id {mat_id} is missing when running add_mc_from_graph")
    _message("running graph clustering now - one iteration no bootstrap")
    edges = {"col1": graph.edges["mc1"], "col2": graph.edges["mc2"], "weight": graph.edges["w"]}
    node_clust = tgs_graph_cover(edges, min_mc_size, get_param("scm_tgs_clust_cool"), get_param("scm_tgs_clust_burn_in"),
                                 n_nodes=len(graph.cell_names), ctx=ctx)
    res = _mc_from_cover(node_clust, mat.cells)
    _message(f"building metacell object, #mc {max(res['mc'].values()) if res['mc'] else 0}")
    _scdb_add("mc", mc_id, res)
    return res


def mcell_coclust_filt_by_k_deg(coc_id, K, alpha):
    """reference R/coclust.r:58-84: boolean filter over the co-cluster rows.  thresh_K[node] = number of distinct
    count levels v (over the whole table) for which the node has more than K incident rows with cnt >= v; a row is
    kept when thresh_K of either endpoint is below cnt * alpha.  Host logic, as in the reference."""
    coc = scdb_coclust(coc_id)
    if coc is None:
        raise McbError(f"MC-ERR: coclust {coc_id} is missing when trying to derive K-deg filter")
    n1, n2, cnt = coc.coclust["node1"], coc.coclust["node2"], coc.coclust["cnt"]
    nnodes = coc.n_samp.size
    levels = np.unique(cnt)                                                 # columns of table(), ascending
    lev = np.searchsorted(levels, cnt)
    deg_wgt = np.zeros((nnodes, levels.size), dtype=np.int64)
    np.add.at(deg_wgt, (n1, lev), 1)
    np.add.at(deg_wgt, (n2, lev), 1)
    deg_cum = np.cumsum(deg_wgt[:, ::-1], axis=1)                          # cumsum(rev(x))
    thresh = (deg_cum > K).sum(axis=1).astype(np.float64)
    present = deg_wgt.sum(axis=1) > 0
    thresh[~present] = np.nan                                               # nodes absent from the table stay NA in R
    with np.errstate(invalid="ignore"):
        return (thresh[n1] < cnt * alpha) | (thresh[n2] < cnt * alpha)


def mcell_mc_from_coclust_balanced(mc_id, coc_id, mat_id, K, min_mc_size, alpha=2, ctx=None):
    """reference R/mc_from_coclust.r:215-263 up to the cluster assignment."""
    coc = scdb_coclust(coc_id)
    if coc is None:
        raise McbError(f"MC-ERR: coclust {coc_id} is missing when running mc from coclust")
    mat = scdb_mat(mat_id)
    if mat is None:
        raise McbError(f"MC-ERR: mat id {mat_id} is missing when running add_mc_from_graph")
    filt = mcell_coclust_filt_by_k_deg(coc_id, K, alpha)
    n1, n2, cnt = (coc.coclust[k][filt] for k in ("node1", "node2", "cnt"))
    _message(f"filtered {filt.size - int(filt.sum())} left with {int(filt.sum())} based on co-cluster imbalance")
    w = cnt / float(cnt.max()) if cnt.size else cnt.astype(np.float64)
    keep = n1 != n2                                                         # :241
    n1, n2, w = n1[keep], n2[keep], w[keep]
    edges = {"col1": np.concatenate([n1, n2]), "col2": np.concatenate([n2, n1]), "weight": np.concatenate([w, w])}   # :242-245
    node_clust = tgs_graph_cover(edges, min_mc_size, get_param("scm_tgs_clust_cool"), get_param("scm_tgs_clust_burn_in"),
                                 n_nodes=coc.n_samp.size, ctx=ctx)
    res = _mc_from_cover(node_clust, mat.cells)
    _message(f"building metacell object, #mc {max(res['mc'].values()) if res['mc'] else 0}")
    _scdb_add("mc", mc_id, res)
    return res
