"""numpy/OpenBLAS restatement of the kNN stage (A5) -- TEST / BASELINE INFRASTRUCTURE ONLY.

tgstat can run its correlation through BLAS (`options(tgs_use.blas = TRUE)`), so the
fair multi-core CPU baseline for tgs_cor_knn (reference call site R/utils.r:119) is an fp64 dgemm in
row blocks followed by a per-row partial selection.  Semantics are those of mc_oracle.c orc_cor_knn:
self forced to rank 1, ties to the lower column index.  Only bench.py's cpu_baseline / --impl
reference legs and tests/ may import this module.  PARITY UNPINNED (see mc_oracle.c).
"""
import numpy as np


def cor_knn_blas(Z, Kc, row_begin=0, row_end=None, block=512):
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    M = Z.shape[0]
    row_end = M if row_end is None else row_end
    kk = min(Kc, M)
    col = np.full((row_end - row_begin, Kc), -1, dtype=np.int32)
    cor = np.zeros((row_end - row_begin, Kc), dtype=np.float64)
    for i0 in range(row_begin, row_end, block):
        i1 = min(row_end, i0 + block)
        C = Z[i0:i1] @ Z.T                                  # dgemm on all cores
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
