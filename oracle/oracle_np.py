"""numpy/OpenBLAS restatement of the kNN stage (A5) -- TEST / BASELINE INFRASTRUCTURE ONLY.

tgstat can run its correlation through BLAS (`options(tgs_use.blas = TRUE)`), so the
fair multi-core CPU baseline for tgs_cor_knn (reference call site R/utils.r:119) is an fp64 dgemm in
row blocks followed by a per-row partial selection (OpenMP, mc_oracle.c orc_select_rows).  Semantics are those of mc_oracle.c orc_cor_knn:
self forced to rank 1, ties to the lower column index.  Only bench.py's cpu_baseline / --impl
reference legs and tests/ may import this module.  PARITY UNPINNED (see mc_oracle.c).
"""
import numpy as np


def cor_knn_blas(Z, Kc, row_begin=0, row_end=None, block=1024, rows=None, timing=None):
    """rows [row_begin, row_end) (or the explicit index list `rows`) against all columns: OpenBLAS dgemm in row
    blocks (all cores) + the OpenMP per-row select of mc_oracle.c (all cores).  `timing` (dict) accumulates the
    seconds spent in each half under "dgemm" / "select"."""
    import time
    from . import oracle as O
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    M = Z.shape[0]
    if rows is None:
        row_end = M if row_end is None else row_end
        rows = np.arange(row_begin, row_end, dtype=np.int64)
    else:
        rows = np.asarray(rows, dtype=np.int64)
    col = np.full((rows.size, Kc), -1, dtype=np.int32)
    cor = np.zeros((rows.size, Kc), dtype=np.float64)
    for b0 in range(0, rows.size, block):
        rb = rows[b0:b0 + block]
        t0 = time.perf_counter()
        contiguous = rb.size > 0 and rb[-1] - rb[0] + 1 == rb.size
        A = Z[rb[0]:rb[-1] + 1] if contiguous else Z[rb]
        C = A @ Z.T                                         # dgemm on all cores
        t1 = time.perf_counter()
        col[b0:b0 + rb.size], cor[b0:b0 + rb.size] = O.select_rows(C, rb, Kc)
        t2 = time.perf_counter()
        if timing is not None:
            timing["dgemm"] = timing.get("dgemm", 0.0) + (t1 - t0)
            timing["select"] = timing.get("select", 0.0) + (t2 - t1)
    return col, cor


def cor_knn_numpy(Z, Kc, row_begin=0, row_end=None, block=512):
    """the same stage with numpy only (argpartition + lexsort; single-threaded selection) -- kept as an independent
    check of orc_select_rows"""
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    M = Z.shape[0]
    row_end = M if row_end is None else row_end
    kk = min(Kc, M)
    col = np.full((row_end - row_begin, Kc), -1, dtype=np.int32)
    cor = np.zeros((row_end - row_begin, Kc), dtype=np.float64)
    for i0 in range(row_begin, row_end, block):
        i1 = min(row_end, i0 + block)
        C = Z[i0:i1] @ Z.T
        rows = np.arange(i0, i1)
        C[np.arange(i1 - i0), rows] = 2.0                   # self sentinel: always rank 1
        if kk < M:
            part = np.argpartition(-C, kk - 1, axis=1)[:, :kk]
        else:
            part = np.broadcast_to(np.arange(M), (i1 - i0, M)).copy()
        vals = np.take_along_axis(C, part, axis=1)
        order = np.lexsort((part, -vals), axis=1)           # (cor desc, col asc)
        part = np.take_along_axis(part, order, axis=1)
        vals = np.take_along_axis(vals, order, axis=1)
        vals[:, 0] = 1.0
        col[i0 - row_begin:i1 - row_begin, :kk] = part
        cor[i0 - row_begin:i1 - row_begin, :kk] = vals
    return col, cor


def standardize_cells(X):
    """A4 without the log transform: X [cells][G] -> (Z, valid); z = (x - mean) / sqrt(sum (x - mean)^2) per cell.
    A cell whose vector is constant has NA correlations in R (R/cgraph_knn_mcnorm.r:25): valid = False."""
    X = np.asarray(X, dtype=np.float64)
    d = X - X.mean(axis=1, keepdims=True)
    s = np.sqrt((d * d).sum(axis=1))
    valid = s > 0
    Z = np.zeros_like(d)
    Z[valid] = d[valid] / s[valid, None]
    return Z, valid


def cor_knn_xy(X, Y, knn):
    """tgs_cor_knn(x, y, knn) with y != x (reference call sites R/atlas.r:164-166,277,475-478,584), fp64:
    per x cell the knn y cells with the highest Pearson correlation, (cor desc, y index asc); no self match.
    Returns col [nx][knn] (-1 where NA), cor [nx][knn] (NaN where NA)."""
    Zx, vx = standardize_cells(X)
    Zy, vy = standardize_cells(Y)
    yi = np.nonzero(vy)[0]
    kk = min(int(knn), yi.size)
    col = np.full((Zx.shape[0], int(knn)), -1, dtype=np.int32)
    cor = np.full((Zx.shape[0], int(knn)), np.nan, dtype=np.float64)
    if kk == 0:
        return col, cor
    C = Zx @ Zy[yi].T
    order = np.lexsort((np.broadcast_to(np.arange(yi.size), C.shape), -C), axis=1)[:, :kk]
    col[vx, :kk] = yi[order[vx]]
    cor[vx, :kk] = np.take_along_axis(C, order, axis=1)[vx]
    return col, cor


def r_round_half_even(v):
    return float(np.rint(v))          # R's round(): IEC 60559 half-to-even, which is numpy's rint


def gene_stats_dense(raw, ds, niche_quantile=0.2, k_std_n=1000.0, sub_cells=None):
    """Literal dense restatement of the per-gene block of scm_gene_stat (reference R/gstats.r:83-87, :102, :113-127,
    :143-151) -- small inputs only.  raw: [genes][cells] counts (all cells); ds: [genes][kept cells] downsampled
    counts (or None).  Returns a dict of unrounded per-gene arrays, one per data-frame column."""
    raw = np.asarray(raw, dtype=np.float64)
    ng, n = raw.shape
    r = int(r_round_half_even(n * niche_quantile))
    out = {}
    out["tot"] = raw.sum(axis=1)
    out["var"] = raw.var(axis=1, ddof=1)
    srt = np.sort(raw, axis=1)
    up = srt[:, n - r:].sum(axis=1) if r > 0 else np.zeros(ng)          # sum(tail(sort(x), n = round(n * q)))
    out["niche_stat"] = (10.0 + up) / (10.0 + out["tot"])
    csum = raw.sum(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(csum > 0, k_std_n / csum, 0.0)
    out["n_mean"] = (raw * scale[None, :]).mean(axis=1)
    out["is_on_count"] = (raw > 0).sum(axis=1).astype(np.float64)
    cells = np.arange(n) if sub_cells is None else np.asarray(sub_cells)
    ct = csum[cells]
    sz = np.full(ng, np.nan)
    for g in range(ng):
        x = raw[g, cells]
        if x.std() > 0 and ct.std() > 0:
            sz[g] = np.corrcoef(x, ct)[0, 1]
    out["sz_cor"] = sz
    if ds is not None:
        ds = np.asarray(ds, dtype=np.float64)
        n_ds = ds.shape[1]
        dsrt = np.sort(ds, axis=1)
        out["ds_top1"] = dsrt[:, n_ds - 1]
        out["ds_top2"] = dsrt[:, n_ds - 2]                                # rowOrderStats(which = n_ds - 1)
        out["ds_top3"] = dsrt[:, n_ds - 3]
        out["ds_mean"] = ds.mean(axis=1)
        out["ds_var"] = ds.var(axis=1, ddof=1)
        out["ds_is_on_count"] = (ds > 0).sum(axis=1).astype(np.float64)
    return out


def rollmedian_fill(x, k, left, right):
    """zoo::rollmedian(x, k, fill = c(left, NA, right)) for odd k: centred running median, ends filled."""
    x = np.asarray(x, dtype=np.float64)
    h = k // 2
    out = np.empty_like(x)
    out[:h] = left
    out[x.size - h:] = right
    for i in range(h, x.size - h):
        out[i] = np.median(x[i - h:i + h + 1])
    return out


def mc_tapply_dense(us, mc_ids):
    """tgs_matrix_tapply(us, mc, f) for f = exp(mean(log(1+y)))-1 and mean(x>0) (reference R/mc.r:185-186, :240), dense
    and literal.  us: [genes][cells] over the covered cells, mc_ids: 0-based metacell per cell.
    Returns (geomean [genes][n_mc], frac_on [genes][n_mc], mc_meansize [n_mc])."""
    us = np.asarray(us, dtype=np.float64)
    n_mc = int(mc_ids.max()) + 1
    geo = np.empty((us.shape[0], n_mc))
    frac = np.empty((us.shape[0], n_mc))
    msz = np.empty(n_mc)
    csum = us.sum(axis=0)
    for k in range(n_mc):
        y = us[:, mc_ids == k]
        geo[:, k] = np.exp(np.log(1.0 + y).mean(axis=1)) - 1.0
        frac[:, k] = (y > 0).mean(axis=1)
        msz[k] = csum[mc_ids == k].mean()
    return geo, frac, msz
