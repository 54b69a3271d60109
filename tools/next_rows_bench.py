"""Timing of the SURVEY 8(f) rows 3 and 4 at a realistic shape (library stage timers, best of 3):
  mcb_mc_gene_tapply  (mc_compute_fp / e_gc / cov_gc): 100,000 cells x 4,000 genes, 1,000 metacells
  mcb_cor_knn_xy      (atlas projection):              20,000 query cells against 100,000 atlas cells, 1,000 genes, knn = 50"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth

ctx = _lib.Context(0)
ctx.set_option("profile", 1)
out = {}

m = synth.make_umi_matrix(100000, 4000, n_types=12, median_umis=3000.0, seed=5)
nnz = int(m.indptr[-1])
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
ids = np.random.default_rng(1).integers(0, 1000, m.ncells).astype(np.int32)
best = None
for rep in range(3):
    ctx.timings_reset()
    t0 = time.perf_counter()
    geo, frac, size, mean = _lib.mc_gene_tapply(ctx, csc, ids, 1000)
    wall = time.perf_counter() - t0
    tm = {k: v[0] for k, v in ctx.timings().items()}
    if best is None or tm["mc_tapply"] < best[0]["mc_tapply"]:
        best = (tm, wall)
k = 7
cells = np.nonzero(ids == k)[0]
dense = np.zeros((m.ngenes, cells.size))
for q, c in enumerate(cells):
    s, e = m.indptr[c], m.indptr[c + 1]
    dense[m.indices[s:e], q] = m.data[s:e]
assert np.allclose(geo[k], np.exp(np.log1p(dense).mean(axis=1)) - 1.0, rtol=1e-10, atol=1e-13)
out["mc_gene_tapply"] = {"cells": m.ncells, "genes": m.ngenes, "nnz": nnz, "metacells": 1000, "kernel_ms": round(best[0]["mc_tapply"], 3),
                         "GBps_alg": round(nnz * 12 / 1e6 / best[0]["mc_tapply"], 1), "call_wall_ms": round(best[1] * 1e3, 1)}
print(out["mc_gene_tapply"], flush=True)
csc.free()

rng = np.random.default_rng(3)
centers = rng.gamma(2.0, 1.0, size=(24, 1000))
def cells_of(n, seed):
    r = np.random.default_rng(seed)
    t = r.integers(0, 24, n)
    return np.log2(1.0 + 7.0 * r.poisson(centers[t] * r.lognormal(0, 0.3, (n, 1)))).astype(np.float64)
X, Y = cells_of(20000, 11), cells_of(100000, 12)
best = None
for rep in range(3):
    ctx.timings_reset()
    t0 = time.perf_counter()
    col, cor = _lib.cor_knn_xy(ctx, X, Y, 50)
    wall = time.perf_counter() - t0
    tm = {k: round(v[0], 3) for k, v in ctx.timings().items()}
    if best is None or wall < best[1]:
        best = (tm, wall)
r = 123
x = X[r] - X[r].mean(); x /= np.sqrt((x * x).sum())
Yc = Y - Y.mean(axis=1, keepdims=True); Yc /= np.sqrt((Yc * Yc).sum(axis=1, keepdims=True))
ref = Yc @ x
assert np.abs(np.sort(ref)[::-1][:50] - cor[r]).max() < 1e-5
flop = 2.0 * 20000 * 100000 * 1000
out["cor_knn_xy"] = {"x_cells": 20000, "y_cells": 100000, "genes": 1000, "knn": 50, "stages_ms": best[0],
                     "alg_TFLOPs": round(flop / 1e9 / best[0]["cor_xy_gemm"], 1), "call_wall_ms": round(best[1] * 1e3, 1)}
print(out["cor_knn_xy"], flush=True)
print(json.dumps(out))
