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
