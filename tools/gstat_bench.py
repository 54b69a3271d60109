"""scm_gene_stat per-gene block at a realistic shape: GPU stages (downsample, bucket, stats) timed with the library's
stage timers, algorithmic bytes vs measured HBM bandwidth, and a numpy/scipy CPU equivalent of the same statistics
(CSR conversion + per-gene reductions + sort-based niche sum) on the host cores for scale."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth

ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ngenes = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
ctx = _lib.Context(0)
t0 = time.time()
m = synth.make_umi_matrix(ncells, ngenes, n_types=12, median_umis=3000.0, seed=5)
print(f"synthetic {ngenes} genes x {ncells} cells, nnz {m.indptr[-1]:,} ({time.time() - t0:.1f} s to generate)", flush=True)
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
csum = csc.col_sums()
n = _lib.downsamp_depth(csum)
ctx.set_option("profile", 1)
best = None
for rep in range(4):
    ctx.timings_reset()
    ds = csc.downsample(float(n), 19, add_non_dsamp=False)
    st = _lib.gene_stats(ctx, csc, ds, 0.2, 1000.0, None)
    tm = {k: v[0] for k, v in ctx.timings().items()}
    ds.free()
    if best is None or sum(tm.values()) < sum(best.values()):
        best = tm
nnz = int(m.indptr[-1])
gpu_ms = sum(best.values())
alg_bytes = nnz * (12 + 12 + 12) + nnz * 8          # scatter read + write, stats read, downsample read/write of x
print("GPU stages ms:", {k: round(v, 3) for k, v in best.items()}, "total", round(gpu_ms, 3))
print(f"gene_bucket+gene_stats: {(best['gene_bucket'] + best['gene_stats']):.3f} ms for {nnz * 36 / 1e9:.2f} GB algorithmic "
      f"-> {nnz * 36 / 1e6 / (best['gene_bucket'] + best['gene_stats']):.0f} GB/s")
# CPU equivalent (scipy CSR + numpy), same statistics except the downsample (timed separately in bench.py)
import scipy.sparse as sp
t0 = time.time()
A = sp.csc_matrix((m.data.astype(np.float64), m.indices, m.indptr), shape=(ngenes, ncells)).tocsr()
tot = np.asarray(A.sum(axis=1)).ravel()
sq = np.asarray(A.multiply(A).sum(axis=1)).ravel()
var = (sq - tot * tot / ncells) / (ncells - 1)
on = np.diff(A.indptr)
scale = 1000.0 / csum
nmean = np.asarray(A.multiply(sp.csr_matrix(scale[None, :])).sum(axis=1)).ravel() / ncells if False else (A @ scale) / ncells
r = int(np.rint(ncells * 0.2))
up = np.empty(ngenes)
for g in range(ngenes):
    v = A.data[A.indptr[g]:A.indptr[g + 1]]
    up[g] = v.sum() if v.size <= r else np.partition(v, v.size - r)[v.size - r:].sum()
szc = (A @ csum.astype(np.float64))
cpu_s = time.time() - t0
print(f"CPU (scipy/numpy, {os.cpu_count()} cores visible, mostly single-threaded): {cpu_s:.2f} s for the raw-matrix statistics")
assert np.array_equal(tot, st["tot"]) and np.allclose(var, st["var"], rtol=1e-9) and np.allclose((10 + up) / (10 + tot), st["niche_stat"], rtol=1e-12)
assert np.allclose(nmean, st["n_mean"], rtol=1e-9)
print(json.dumps({"gene_stats": {"genes": ngenes, "cells": ncells, "nnz": nnz, "gpu_ms": round(gpu_ms, 3),
                                 "stages_ms": {k: round(v, 3) for k, v in best.items()}, "cpu_equiv_s": round(cpu_s, 2)}}))
